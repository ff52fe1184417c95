// Fill-reducing ordering: approximate minimum degree on a quotient graph.
//
// Replaces what the reference obtains from its un-vendored dependency chain
// MadNLP.LDLSolver -> LDLFactorizations.ldl_analyze -> AMD.jl (call site
// /root/reference/src/KKT/normalkkt.jl:99-101; solver choice
// /root/reference/scripts/benchmarks_cpu.jl:36).  Written from the published
// algorithm (Amestoy, Davis, Duff, "An approximate minimum degree ordering
// algorithm", SIMAX 17(4), 1996): quotient graph stored in one integer
// workspace, approximate external degrees from the |Le \ Lme| trick, element
// absorption (incl. aggressive), mass elimination, supervariable detection by
// hashing, dense rows ordered last, final ordering by a postorder of the
// assembly tree.  Bit-parity with SuiteSparse AMD is NOT claimed (the
// reference's permutation cannot be produced in this environment).
#include "symbolic.hpp"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace madipm {

namespace {

constexpr int32_t kNone = -1;
inline int32_t flip(int32_t i) { return -i - 2; }

struct QuotientGraph {
    int32_t n = 0;
    std::vector<int32_t> start;    // list start in work[]; for absorbed nodes: flip(parent)
    std::vector<int32_t> len;      // list length
    std::vector<int32_t> elen;     // #elements at the front of a variable's list; -1 absorbed variable; <-1 element (flip(size))
    std::vector<int32_t> weight;   // supervariable size; 0 = non-principal; negative while in the pivot element
    std::vector<int32_t> degree;   // approximate external degree
    std::vector<int32_t> mark;     // the w[] array of the algorithm
    std::vector<int32_t> work;     // adjacency workspace
    int64_t free_at = 0;
    // degree lists
    std::vector<int32_t> dhead, dnext, dprev;
    // hash buckets for supervariable detection
    std::vector<int32_t> hhead, hnext, hkey;
};

// Compact the workspace: every live list (start >= 0, len > 0) is moved to the
// front, keeping relative order.  Only called between pivots.
void garbage_collect(QuotientGraph& g) {
    const int32_t n = g.n;
    std::vector<std::pair<int32_t, int32_t>> live;
    live.reserve(n);
    for (int32_t i = 0; i < n; ++i)
        if (g.start[i] >= 0 && g.len[i] > 0) live.emplace_back(g.start[i], i);
    std::sort(live.begin(), live.end());
    int64_t dst = 0;
    for (auto& sl : live) {
        int32_t i = sl.second;
        int64_t src = sl.first;
        g.start[i] = static_cast<int32_t>(dst);
        for (int32_t k = 0; k < g.len[i]; ++k) g.work[dst++] = g.work[src++];
    }
    g.free_at = dst;
}

inline void dlist_remove(QuotientGraph& g, int32_t i) {
    int32_t nx = g.dnext[i], pv = g.dprev[i];
    if (nx != kNone) g.dprev[nx] = pv;
    if (pv != kNone) g.dnext[pv] = nx;
    else g.dhead[g.degree[i]] = nx;
}

inline void dlist_insert(QuotientGraph& g, int32_t i, int32_t d) {
    int32_t h = g.dhead[d];
    g.dnext[i] = h;
    g.dprev[i] = kNone;
    if (h != kNone) g.dprev[h] = i;
    g.dhead[d] = i;
}

int32_t reset_marks(int32_t wflg, int32_t wbig, std::vector<int32_t>& mark) {
    if (wflg < 2 || wflg >= wbig) {
        for (auto& m : mark) if (m != 0) m = 1;
        wflg = 2;
    }
    return wflg;
}

}  // namespace

// Input: symmetric pattern as CSC of the strict lower OR any triangle/both (the
// routine symmetrises and drops the diagonal).  Output: perm[new] = old.
void amd_order(int32_t n, const int64_t* colptr, const int32_t* rowval, std::vector<int32_t>& perm,
               double dense_factor) {
    perm.assign(n, 0);
    if (n == 0) return;

    // ---- symmetrised adjacency, duplicates removed, no diagonal
    std::vector<int64_t> cnt(n + 1, 0);
    for (int32_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            int32_t i = rowval[p];
            if (i == j) continue;
            ++cnt[i + 1];
            ++cnt[j + 1];
        }
    for (int32_t i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    std::vector<int32_t> adj(cnt[n]);
    {
        std::vector<int64_t> at(cnt.begin(), cnt.end() - 1);
        for (int32_t j = 0; j < n; ++j)
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                int32_t i = rowval[p];
                if (i == j) continue;
                adj[at[i]++] = j;
                adj[at[j]++] = i;
            }
    }
    std::vector<int64_t> aptr(n + 1, 0);
    {
        std::vector<int32_t> seen(n, kNone);
        int64_t out = 0;
        for (int32_t i = 0; i < n; ++i) {
            int64_t b = out;
            for (int64_t p = cnt[i]; p < cnt[i + 1]; ++p) {
                int32_t j = adj[p];
                if (seen[j] != i) { seen[j] = i; adj[out++] = j; }
            }
            aptr[i] = b;
        }
        aptr[n] = out;
        // aptr currently holds begins; fine as [aptr[i], aptr[i+1])
    }
    const int64_t nz = aptr[n];
    if (nz + n > INT32_MAX / 2) throw std::runtime_error("amd_order: pattern too large for 32-bit workspace");

    QuotientGraph g;
    g.n = n;
    g.start.assign(n, 0);
    g.len.assign(n, 0);
    g.elen.assign(n, 0);
    g.weight.assign(n, 1);
    g.degree.assign(n, 0);
    g.mark.assign(n, 1);
    g.dhead.assign(n + 1, kNone);
    g.dnext.assign(n, kNone);
    g.dprev.assign(n, kNone);
    g.hhead.assign(n, kNone);
    g.hnext.assign(n, kNone);
    g.hkey.assign(n, 0);
    const int64_t wlen = nz + nz / 5 + 2 * static_cast<int64_t>(n) + 16;
    g.work.assign(wlen, 0);
    for (int32_t i = 0; i < n; ++i) {
        g.start[i] = static_cast<int32_t>(aptr[i]);
        g.len[i] = static_cast<int32_t>(aptr[i + 1] - aptr[i]);
        g.degree[i] = g.len[i];
    }
    std::copy(adj.begin(), adj.begin() + nz, g.work.begin());
    g.free_at = nz;
    adj.clear();
    adj.shrink_to_fit();

    int32_t dense_cut;
    if (dense_factor < 0) dense_cut = n - 2;
    else dense_cut = static_cast<int32_t>(std::min<double>(n - 2.0, std::max(16.0, dense_factor * std::sqrt(static_cast<double>(n)))));
    if (dense_cut < 1) dense_cut = 1;

    int32_t nel = 0, ndense = 0;
    std::vector<char> kind(n, 0);   // 0 ordinary, 2 isolated, 3 dense
    for (int32_t i = 0; i < n; ++i) {
        int32_t d = g.degree[i];
        if (d == 0) {
            kind[i] = 2;
            // isolated: eliminate right away as an empty element
            g.elen[i] = flip(1);
            ++nel;
            g.start[i] = kNone;
            g.mark[i] = 0;
        } else if (d > dense_cut) {
            // dense: ordered last, removed from the graph
            kind[i] = 3;
            ++ndense;
            g.weight[i] = 0;
            g.elen[i] = kNone;
            ++nel;
            g.start[i] = kNone;
        } else {
            dlist_insert(g, i, d);
        }
    }

    const int32_t wbig = INT32_MAX - n;
    int32_t wflg = 2;
    int32_t mindeg = 0;
    int32_t lemax = 0;
    std::vector<int32_t> elim_order;   // principal pivots in elimination order
    elim_order.reserve(n);

    while (nel < n) {
        // ---- pivot of minimum approximate degree
        int32_t me = kNone;
        for (; mindeg < n; ++mindeg) {
            me = g.dhead[mindeg];
            if (me != kNone) break;
        }
        if (me == kNone) throw std::runtime_error("amd_order: degree lists exhausted");
        dlist_remove(g, me);
        elim_order.push_back(me);
        const int32_t elenme = g.elen[me];
        int32_t nvpiv = g.weight[me];
        nel += nvpiv;

        // ---- build the new element Lme
        g.weight[me] = -nvpiv;
        int32_t degme = 0;
        int64_t pme1, pme2;
        if (elenme == 0) {
            // in place: only variables in the list
            pme1 = g.start[me];
            pme2 = pme1 - 1;
            for (int64_t p = pme1; p < pme1 + g.len[me]; ++p) {
                int32_t i = g.work[p];
                int32_t nvi = g.weight[i];
                if (nvi > 0) {
                    degme += nvi;
                    g.weight[i] = -nvi;
                    g.work[++pme2] = i;
                    dlist_remove(g, i);
                }
            }
        } else {
            // room for the new element: at most the sum of the absorbed lists
            {
                int64_t need = g.len[me] - elenme;
                for (int64_t q = g.start[me]; q < g.start[me] + elenme; ++q) need += g.len[g.work[q]];
                need = std::min<int64_t>(need, n);
                if (g.free_at + need > static_cast<int64_t>(g.work.size())) {
                    garbage_collect(g);
                    if (g.free_at + need > static_cast<int64_t>(g.work.size()))
                        g.work.resize(g.free_at + need + g.work.size() / 4 + 16);
                }
            }
            int64_t p = g.start[me];
            int64_t es = g.free_at, ee = g.free_at;
            const int32_t slenme = g.len[me] - elenme;
            for (int32_t k = 0; k <= elenme; ++k) {
                int32_t e;
                int64_t pj;
                int32_t ln;
                if (k == elenme) { e = me; pj = p; ln = slenme; }
                else { e = g.work[p++]; pj = g.start[e]; ln = g.len[e]; }
                for (int32_t t = 0; t < ln; ++t) {
                    int32_t i = g.work[pj++];
                    int32_t nvi = g.weight[i];
                    if (nvi > 0) {
                        degme += nvi;
                        g.weight[i] = -nvi;
                        g.work[ee++] = i;
                        dlist_remove(g, i);
                    }
                }
                if (e != me) {
                    g.start[e] = flip(me);
                    g.mark[e] = 0;
                }
            }
            pme1 = es;
            pme2 = ee - 1;
            g.free_at = ee;
        }
        g.degree[me] = degme;
        g.start[me] = static_cast<int32_t>(pme1);
        g.len[me] = static_cast<int32_t>(pme2 - pme1 + 1);
        g.elen[me] = flip(nvpiv + degme);

        wflg = reset_marks(wflg, wbig, g.mark);

        // ---- |Le \ Lme| for every element e adjacent to a variable of Lme
        for (int64_t pme = pme1; pme <= pme2; ++pme) {
            int32_t i = g.work[pme];
            int32_t eln = g.elen[i];
            if (eln > 0) {
                int32_t nvi = -g.weight[i];
                int32_t wnvi = wflg - nvi;
                for (int64_t p = g.start[i]; p < g.start[i] + eln; ++p) {
                    int32_t e = g.work[p];
                    int32_t we = g.mark[e];
                    if (we >= wflg) we -= nvi;
                    else if (we != 0) we = g.degree[e] + wnvi;
                    g.mark[e] = we;
                }
            }
        }

        // ---- degree update, absorption, hashing
        for (int64_t pme = pme1; pme <= pme2; ++pme) {
            int32_t i = g.work[pme];
            int64_t p1 = g.start[i];
            int64_t p2 = p1 + g.elen[i] - 1;
            int64_t pn = p1;
            uint32_t hash = 0;
            int32_t deg = 0;
            for (int64_t p = p1; p <= p2; ++p) {
                int32_t e = g.work[p];
                int32_t we = g.mark[e];
                if (we != 0) {
                    int32_t dext = we - wflg;
                    if (dext > 0) {
                        deg += dext;
                        g.work[pn++] = e;
                        hash += static_cast<uint32_t>(e);
                    } else {
                        g.start[e] = flip(me);   // aggressive absorption
                        g.mark[e] = 0;
                    }
                }
            }
            g.elen[i] = static_cast<int32_t>(pn - p1 + 1);
            int64_t p3 = pn;
            int64_t p4 = p1 + g.len[i];
            for (int64_t p = p2 + 1; p < p4; ++p) {
                int32_t j = g.work[p];
                int32_t nvj = g.weight[j];
                if (nvj > 0) {
                    deg += nvj;
                    g.work[pn++] = j;
                    hash += static_cast<uint32_t>(j);
                }
            }
            if (g.elen[i] == 1 && p3 == pn) {
                // mass elimination: i is adjacent to me only
                g.start[i] = flip(me);
                int32_t nvi = -g.weight[i];
                degme -= nvi;
                nvpiv += nvi;
                nel += nvi;
                g.weight[i] = 0;
                g.elen[i] = kNone;
            } else {
                g.degree[i] = std::min(g.degree[i], deg);
                // make room for `me` at the head of the element list
                g.work[pn] = g.work[p3];
                g.work[p3] = g.work[p1];
                g.work[p1] = me;
                g.len[i] = static_cast<int32_t>(pn - p1 + 1);
                int32_t h = static_cast<int32_t>(hash % static_cast<uint32_t>(n));
                g.hkey[i] = h;
                g.hnext[i] = g.hhead[h];
                g.hhead[h] = i;
            }
        }
        g.degree[me] = degme;
        lemax = std::max(lemax, degme);
        wflg += lemax;
        wflg = reset_marks(wflg, wbig, g.mark);

        // ---- supervariable detection within hash buckets
        for (int64_t pme = pme1; pme <= pme2; ++pme) {
            int32_t iv = g.work[pme];
            if (g.weight[iv] >= 0) continue;
            int32_t h = g.hkey[iv];
            int32_t i = g.hhead[h];
            g.hhead[h] = kNone;             // consume the bucket once
            while (i != kNone && g.hnext[i] != kNone) {
                const int32_t ln = g.len[i];
                const int32_t eln = g.elen[i];
                for (int64_t p = g.start[i] + 1; p < g.start[i] + ln; ++p) g.mark[g.work[p]] = wflg;
                int32_t jlast = i;
                int32_t j = g.hnext[i];
                while (j != kNone) {
                    bool same = (g.len[j] == ln) && (g.elen[j] == eln);
                    for (int64_t p = g.start[j] + 1; same && p < g.start[j] + ln; ++p)
                        if (g.mark[g.work[p]] != wflg) same = false;
                    if (same) {
                        g.start[j] = flip(i);
                        g.weight[i] += g.weight[j];   // both negative here
                        g.weight[j] = 0;
                        g.elen[j] = kNone;
                        j = g.hnext[j];
                        g.hnext[jlast] = j;
                    } else {
                        jlast = j;
                        j = g.hnext[j];
                    }
                }
                ++wflg;
                i = g.hnext[i];
            }
        }

        // ---- put the surviving variables back in the degree lists
        int64_t p = pme1;
        const int32_t nleft = n - nel;
        for (int64_t pme = pme1; pme <= pme2; ++pme) {
            int32_t i = g.work[pme];
            int32_t nvi = -g.weight[i];
            if (nvi > 0) {
                g.weight[i] = nvi;
                int32_t deg = g.degree[i] + degme - nvi;
                deg = std::min(deg, nleft - nvi);
                if (deg < 0) deg = 0;
                g.degree[i] = deg;
                dlist_insert(g, i, deg);
                mindeg = std::min(mindeg, deg);
                g.work[p++] = i;
                g.hnext[i] = kNone;
            }
        }
        g.weight[me] = nvpiv;
        g.len[me] = static_cast<int32_t>(p - pme1);
        if (g.len[me] == 0) {
            g.start[me] = kNone;
            g.mark[me] = 0;
        }
        if (elenme != 0) g.free_at = p;
    }

    // ---- assembly tree -> permutation.
    // parent of an absorbed node i is flip(start[i]); kind[] says what each
    // node ended up as.
    std::vector<int32_t> tparent(n, kNone);
    for (int32_t i = 0; i < n; ++i)
        if (g.start[i] < kNone) tparent[i] = flip(g.start[i]);
    std::vector<char> is_elem(n, 0);
    for (int32_t e : elim_order) is_elem[e] = 1;
    for (int32_t i = 0; i < n; ++i)
        if (kind[i] == 2) is_elem[i] = 2;   // isolated nodes eliminated up front

    // find, for every variable, the element it was finally absorbed into
    auto elem_of = [&](int32_t i) {
        int32_t e = i;
        while (!is_elem[e] && tparent[e] != kNone) e = tparent[e];
        return e;
    };
    // element tree: parent element of an element
    std::vector<int32_t> eparent(n, kNone);
    for (int32_t e : elim_order) {
        int32_t q = tparent[e];
        while (q != kNone && !is_elem[q]) q = tparent[q];
        eparent[e] = q;
    }
    // children lists in elimination order, then iterative postorder
    std::vector<int32_t> order_rank(n, kNone);
    for (size_t k = 0; k < elim_order.size(); ++k) order_rank[elim_order[k]] = static_cast<int32_t>(k);
    std::vector<int32_t> chead(n, kNone), cnext(n, kNone);
    std::vector<int32_t> roots;
    for (int64_t k = static_cast<int64_t>(elim_order.size()) - 1; k >= 0; --k) {
        int32_t e = elim_order[k];
        int32_t q = eparent[e];
        if (q == kNone) roots.push_back(e);
        else { cnext[e] = chead[q]; chead[q] = e; }
    }
    std::reverse(roots.begin(), roots.end());
    // members (non-principal variables) of each element, in index order
    std::vector<int32_t> mhead(n, kNone), mnext(n, kNone);
    std::vector<int32_t> dense_nodes, isolated;
    for (int32_t i = n - 1; i >= 0; --i) {
        if (is_elem[i] == 1) continue;
        if (is_elem[i] == 2) { isolated.push_back(i); continue; }
        int32_t e = elem_of(i);
        if (kind[i] == 3 || e == i || !is_elem[e]) { dense_nodes.push_back(i); continue; }
        mnext[i] = mhead[e];
        mhead[e] = i;
    }
    std::reverse(dense_nodes.begin(), dense_nodes.end());
    std::reverse(isolated.begin(), isolated.end());

    int32_t k = 0;
    for (int32_t i : isolated) perm[k++] = i;
    std::vector<int32_t> stack;
    std::vector<int32_t> iter(n, kNone);
    for (int32_t r : roots) {
        stack.push_back(r);
        iter[r] = chead[r];
        while (!stack.empty()) {
            int32_t e = stack.back();
            int32_t c = iter[e];
            if (c != kNone) {
                iter[e] = cnext[c];
                stack.push_back(c);
                iter[c] = chead[c];
            } else {
                stack.pop_back();
                perm[k++] = e;
                for (int32_t v = mhead[e]; v != kNone; v = mnext[v]) perm[k++] = v;
            }
        }
    }
    for (int32_t i : dense_nodes) perm[k++] = i;
    if (k != n) throw std::runtime_error("amd_order: internal error, permutation incomplete");
    (void)ndense;
}

}  // namespace madipm
