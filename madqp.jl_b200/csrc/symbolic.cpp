// Host-side symbolic analysis.  See symbolic.hpp for the reference citations.
#include "symbolic.hpp"

#include <atomic>
#include <chrono>
#include <exception>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <climits>
#include <cstring>
#include <numeric>
#include <stdexcept>

namespace madipm {

// ------------------------------------------------------------------------------------------
// S1  stable counting sort, /root/reference/src/utils.jl:148-191 (0-based restatement)
// ------------------------------------------------------------------------------------------
void coo_to_csr(int32_t n_rows, int64_t nnz, const int32_t* Ai, const int32_t* Aj,
                std::vector<int64_t>& Bp, std::vector<int32_t>& Bj, std::vector<int64_t>& order) {
    Bp.assign(static_cast<size_t>(n_rows) + 1, 0);
    Bj.assign(nnz, 0);
    order.assign(nnz, 0);
    for (int64_t k = 0; k < nnz; ++k) {
        if (Ai[k] < 0 || Ai[k] >= n_rows) throw std::runtime_error("coo_to_csr: row index out of range");
        ++Bp[Ai[k] + 1];
    }
    for (int32_t i = 0; i < n_rows; ++i) Bp[i + 1] += Bp[i];
    std::vector<int64_t> at(Bp.begin(), Bp.end() - 1);
    for (int64_t k = 0; k < nnz; ++k) {
        int64_t d = at[Ai[k]]++;
        Bj[d] = Aj[k];
        order[d] = k;
    }
}

// ------------------------------------------------------------------------------------------
// S3  pattern of tril(A A'), same output as /root/reference/src/utils.jl:199-264
// ------------------------------------------------------------------------------------------
void build_normal_pattern(int32_t n_rows, int32_t n_cols, const int64_t* Ap, const int32_t* Aj,
                          std::vector<int64_t>& Cp, std::vector<int32_t>& Cj) {
    // column lists of A (rows ascending because rows are scanned in order)
    std::vector<int64_t> cptr(static_cast<size_t>(n_cols) + 1, 0);
    const int64_t nnz = Ap[n_rows];
    for (int64_t p = 0; p < nnz; ++p) ++cptr[Aj[p] + 1];
    for (int32_t k = 0; k < n_cols; ++k) cptr[k + 1] += cptr[k];
    std::vector<int32_t> crow(nnz);
    {
        std::vector<int64_t> at(cptr.begin(), cptr.end() - 1);
        for (int32_t i = 0; i < n_rows; ++i)
            for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) crow[at[Aj[p]]++] = i;
    }
    std::vector<int64_t> live(cptr.begin(), cptr.end() - 1);   // first entry with row >= current i
    std::vector<int32_t> seen(n_rows, -1);
    Cp.assign(static_cast<size_t>(n_rows) + 1, 0);
    Cj.clear();
    std::vector<int32_t> buf;
    for (int32_t i = 0; i < n_rows; ++i) {
        buf.clear();
        for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
            int32_t k = Aj[p];
            int64_t q = live[k];
            while (q < cptr[k + 1] && crow[q] < i) ++q;
            live[k] = q;
            for (; q < cptr[k + 1]; ++q) {
                int32_t j = crow[q];
                if (seen[j] != i) { seen[j] = i; buf.push_back(j); }
            }
        }
        std::sort(buf.begin(), buf.end());
        Cj.insert(Cj.end(), buf.begin(), buf.end());
        Cp[i + 1] = static_cast<int64_t>(Cj.size());
    }
}

namespace {

void transpose_pattern(int32_t n_rows, int32_t n_cols, const std::vector<int64_t>& Ap, const std::vector<int32_t>& Aj,
                       std::vector<int64_t>& Tp, std::vector<int32_t>& Tj, std::vector<int64_t>& Tsrc) {
    const int64_t nnz = Ap[n_rows];
    Tp.assign(static_cast<size_t>(n_cols) + 1, 0);
    Tj.assign(nnz, 0);
    Tsrc.assign(nnz, 0);
    for (int64_t p = 0; p < nnz; ++p) ++Tp[Aj[p] + 1];
    for (int32_t k = 0; k < n_cols; ++k) Tp[k + 1] += Tp[k];
    std::vector<int64_t> at(Tp.begin(), Tp.end() - 1);
    for (int32_t i = 0; i < n_rows; ++i)
        for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) {
            int64_t d = at[Aj[p]]++;
            Tj[d] = i;
            Tsrc[d] = p;
        }
}

void build_jacobian(int32_t nx, int32_t m, int64_t nnzj, const int32_t* J_I, const int32_t* J_J,
                    int32_t n_slack, const int64_t* ind_ineq, KktPattern& out) {
    out.nx = nx;
    out.n_slack = n_slack;
    out.n_tot = nx + n_slack;
    out.m = m;
    out.nnzj = nnzj;
    // COO order [jac | slack], /root/reference/src/KKT/normalkkt.jl:67-77
    out.jac_I.resize(nnzj + n_slack);
    out.jac_J.resize(nnzj + n_slack);
    for (int64_t k = 0; k < nnzj; ++k) {
        int32_t i = J_I[k] - 1, j = J_J[k] - 1;
        if (i < 0 || i >= m || j < 0 || j >= nx) throw std::runtime_error("jacobian index out of range");
        out.jac_I[k] = i;
        out.jac_J[k] = j;
    }
    for (int32_t s = 0; s < n_slack; ++s) {
        int64_t r = ind_ineq[s] - 1;
        if (r < 0 || r >= m) throw std::runtime_error("ind_ineq out of range");
        out.jac_I[nnzj + s] = static_cast<int32_t>(r);
        out.jac_J[nnzj + s] = nx + s;
    }
    coo_to_csr(m, nnzj + n_slack, out.jac_I.data(), out.jac_J.data(), out.A_rowptr, out.A_colind, out.A_csr_order);
    transpose_pattern(m, out.n_tot, out.A_rowptr, out.A_colind, out.AT_rowptr, out.AT_colind, out.AT_src);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// S4  K2 augmented system: COO block order  pr_diag | hess | jac | slack | du_diag
// (MadNLP SparseKKTSystem, un-vendored; block order corroborated by
// /root/reference/scripts/cuda_wrapper.jl:95-107), CSC = sorted + duplicate-merged.
// ------------------------------------------------------------------------------------------
void build_k2_pattern(int32_t nx, int32_t m, int64_t nnzh, const int32_t* H_I, const int32_t* H_J,
                      int64_t nnzj, const int32_t* J_I, const int32_t* J_J,
                      int32_t n_slack, const int64_t* ind_ineq, KktPattern& out) {
    out = KktPattern();
    out.kind = 0;
    build_jacobian(nx, m, nnzj, J_I, J_J, n_slack, ind_ineq, out);
    const int32_t n_tot = out.n_tot;
    out.nnzh = nnzh;
    out.N = n_tot + m;
    out.hess_I.resize(nnzh);
    out.hess_J.resize(nnzh);
    for (int64_t k = 0; k < nnzh; ++k) {
        int32_t i = H_I[k] - 1, j = H_J[k] - 1;
        if (i < 0 || i >= nx || j < 0 || j >= nx) throw std::runtime_error("hessian index out of range");
        if (i < j) std::swap(i, j);   // force_lower_triangular!
        out.hess_I[k] = i;
        out.hess_J[k] = j;
    }
    // full symmetric CSR of Q for the Qx product (/root/reference/scripts/cuda_wrapper.jl:60)
    {
        std::vector<int64_t>& qp = out.Q_rowptr;
        qp.assign(static_cast<size_t>(n_tot) + 1, 0);
        for (int64_t k = 0; k < nnzh; ++k) {
            ++qp[out.hess_I[k] + 1];
            if (out.hess_I[k] != out.hess_J[k]) ++qp[out.hess_J[k] + 1];
        }
        for (int32_t i = 0; i < n_tot; ++i) qp[i + 1] += qp[i];
        out.Q_colind.assign(qp[n_tot], 0);
        out.Q_src.assign(qp[n_tot], 0);
        std::vector<int64_t> at(qp.begin(), qp.end() - 1);
        for (int64_t k = 0; k < nnzh; ++k) {
            int32_t i = out.hess_I[k], j = out.hess_J[k];
            int64_t d = at[i]++;
            out.Q_colind[d] = j;
            out.Q_src[d] = k;
            if (i != j) {
                d = at[j]++;
                out.Q_colind[d] = i;
                out.Q_src[d] = k;
            }
        }
    }
    // aug COO
    const int64_t len = static_cast<int64_t>(n_tot) + nnzh + nnzj + n_slack + m;
    std::vector<int32_t> I(len), J(len);
    int64_t k = 0;
    for (int32_t i = 0; i < n_tot; ++i, ++k) { I[k] = i; J[k] = i; }
    for (int64_t t = 0; t < nnzh; ++t, ++k) { I[k] = out.hess_I[t]; J[k] = out.hess_J[t]; }
    for (int64_t t = 0; t < nnzj + n_slack; ++t, ++k) { I[k] = out.jac_I[t] + n_tot; J[k] = out.jac_J[t]; }
    for (int32_t i = 0; i < m; ++i, ++k) { I[k] = n_tot + i; J[k] = n_tot + i; }
    // CSC with duplicates merged: counting sort by column, then sort rows inside the column
    const int32_t N = out.N;
    std::vector<int64_t> cp(static_cast<size_t>(N) + 1, 0);
    for (int64_t t = 0; t < len; ++t) ++cp[J[t] + 1];
    for (int32_t j = 0; j < N; ++j) cp[j + 1] += cp[j];
    std::vector<int64_t> slot(len);
    {
        std::vector<int64_t> at(cp.begin(), cp.end() - 1);
        for (int64_t t = 0; t < len; ++t) slot[at[J[t]]++] = t;
    }
    out.kkt_colptr.assign(static_cast<size_t>(N) + 1, 0);
    out.kkt_rowval.clear();
    out.kkt_rowval.reserve(len);
    out.aug_csc_map.assign(len, -1);
    out.diag_slot.assign(N, -1);
    for (int32_t j = 0; j < N; ++j) {
        auto b = slot.begin() + cp[j], e = slot.begin() + cp[j + 1];
        std::sort(b, e, [&](int64_t a, int64_t c) { return I[a] < I[c] || (I[a] == I[c] && a < c); });
        int32_t last = -1;
        for (auto it = b; it != e; ++it) {
            int32_t r = I[*it];
            if (r != last) {
                out.kkt_rowval.push_back(r);
                last = r;
                if (r == j) out.diag_slot[j] = static_cast<int64_t>(out.kkt_rowval.size()) - 1;
            }
            out.aug_csc_map[*it] = static_cast<int64_t>(out.kkt_rowval.size()) - 1;
        }
        out.kkt_colptr[j + 1] = static_cast<int64_t>(out.kkt_rowval.size());
    }
}

// ------------------------------------------------------------------------------------------
// normal equations:  aug_com = tril(A A') (/root/reference/src/KKT/normalkkt.jl:94-97) plus
// the product lists that make K2b (utils.jl:266-298) a map-driven kernel.
// ------------------------------------------------------------------------------------------
void build_normal_kkt_pattern(int32_t nx, int32_t m, int64_t nnzj, const int32_t* J_I, const int32_t* J_J,
                              int32_t n_slack, const int64_t* ind_ineq, KktPattern& out) {
    out = KktPattern();
    out.kind = 1;
    build_jacobian(nx, m, nnzj, J_I, J_J, n_slack, ind_ineq, out);
    out.N = m;
    out.Q_rowptr.assign(static_cast<size_t>(out.n_tot) + 1, 0);
    build_normal_pattern(m, out.n_tot, out.A_rowptr.data(), out.A_colind.data(), out.kkt_colptr, out.kkt_rowval);
    const int64_t nnzC = out.kkt_colptr[m];
    out.prod_ptr.assign(static_cast<size_t>(nnzC) + 1, 0);
    std::vector<int32_t> where(out.n_tot, -1);
    for (int32_t i = 0; i < m; ++i) {
        for (int64_t p = out.A_rowptr[i]; p < out.A_rowptr[i + 1]; ++p) where[out.A_colind[p]] = static_cast<int32_t>(p);
        for (int64_t c = out.kkt_colptr[i]; c < out.kkt_colptr[i + 1]; ++c) {
            int32_t j = out.kkt_rowval[c];
            // summation order = order of row j's entries, as in utils.jl:287-290
            for (int64_t d = out.A_rowptr[j]; d < out.A_rowptr[j + 1]; ++d) {
                int32_t k = out.A_colind[d];
                if (where[k] >= 0) {
                    out.prod_a.push_back(where[k]);
                    out.prod_b.push_back(static_cast<int32_t>(d));
                    out.prod_k.push_back(k);
                }
            }
            out.prod_ptr[c + 1] = static_cast<int64_t>(out.prod_a.size());
        }
        for (int64_t p = out.A_rowptr[i]; p < out.A_rowptr[i + 1]; ++p) where[out.A_colind[p]] = -1;
    }
}

// ------------------------------------------------------------------------------------------
// S5  elimination tree, postorder, column counts, supernodes, maps, schedule
// ------------------------------------------------------------------------------------------
namespace {

// strict-lower structure of P A P' as row lists (cols < row) and column lists (rows > col)
struct PermutedPattern {
    std::vector<int64_t> rptr, cptr;
    std::vector<int32_t> rcol, crow;
};

void permute_pattern(int32_t n, const int64_t* colptr, const int32_t* rowval, const std::vector<int32_t>& iperm,
                     PermutedPattern& pp) {
    pp.rptr.assign(static_cast<size_t>(n) + 1, 0);
    pp.cptr.assign(static_cast<size_t>(n) + 1, 0);
    for (int32_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            int32_t i = rowval[p];
            if (i == j) continue;
            int32_t a = iperm[i], b = iperm[j];
            int32_t r = std::max(a, b), c = std::min(a, b);
            ++pp.rptr[r + 1];
            ++pp.cptr[c + 1];
        }
    for (int32_t i = 0; i < n; ++i) { pp.rptr[i + 1] += pp.rptr[i]; pp.cptr[i + 1] += pp.cptr[i]; }
    pp.rcol.assign(pp.rptr[n], 0);
    pp.crow.assign(pp.cptr[n], 0);
    std::vector<int64_t> ra(pp.rptr.begin(), pp.rptr.end() - 1), ca(pp.cptr.begin(), pp.cptr.end() - 1);
    for (int32_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            int32_t i = rowval[p];
            if (i == j) continue;
            int32_t a = iperm[i], b = iperm[j];
            int32_t r = std::max(a, b), c = std::min(a, b);
            pp.rcol[ra[r]++] = c;
            pp.crow[ca[c]++] = r;
        }
}

// Liu's algorithm with path compression
void etree(int32_t n, const PermutedPattern& pp, std::vector<int32_t>& parent) {
    parent.assign(n, -1);
    std::vector<int32_t> anc(n, -1);
    for (int32_t k = 0; k < n; ++k)
        for (int64_t p = pp.rptr[k]; p < pp.rptr[k + 1]; ++p) {
            int32_t i = pp.rcol[p];
            while (i != -1 && i < k) {
                int32_t nxt = anc[i];
                anc[i] = k;
                if (nxt == -1) parent[i] = k;
                i = nxt;
            }
        }
}

// children visited in the order given by `key` ascending (ties by index); returns post[k] = node
void postorder(int32_t n, const std::vector<int32_t>& parent, const std::vector<int32_t>* key,
               std::vector<int32_t>& post) {
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    if (key) std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return (*key)[a] < (*key)[b]; });
    // build child lists so that iteration yields ascending key: push in descending order
    std::vector<int32_t> head(n, -1), next(n, -1);
    std::vector<int32_t> roots;
    for (int32_t t = n - 1; t >= 0; --t) {
        int32_t v = order[t];
        if (parent[v] == -1) roots.push_back(v);
        else { next[v] = head[parent[v]]; head[parent[v]] = v; }
    }
    std::reverse(roots.begin(), roots.end());
    post.clear();
    post.reserve(n);
    std::vector<int32_t> stack;
    for (int32_t r : roots) {
        stack.push_back(r);
        while (!stack.empty()) {
            int32_t v = stack.back();
            int32_t c = head[v];
            if (c != -1) { head[v] = next[c]; stack.push_back(c); }
            else { stack.pop_back(); post.push_back(v); }
        }
    }
}

// Column counts of L (Gilbert, Ng, Peyton 1994), skeleton-matrix formulation.
void column_counts(int32_t n, const PermutedPattern& pp, const std::vector<int32_t>& parent,
                   const std::vector<int32_t>& post, std::vector<int32_t>& count) {
    std::vector<int32_t> first(n, -1), maxfirst(n, -1), prevleaf(n, -1), anc(n);
    std::vector<int64_t> delta(n, 0);
    std::iota(anc.begin(), anc.end(), 0);
    for (int32_t k = 0; k < n; ++k) {
        int32_t j = post[k];
        delta[j] = (first[j] == -1) ? 1 : 0;
        for (; j != -1 && first[j] == -1; j = parent[j]) first[j] = k;
    }
    for (int32_t k = 0; k < n; ++k) {
        int32_t j = post[k];
        if (parent[j] != -1) --delta[parent[j]];
        for (int64_t p = pp.cptr[j]; p < pp.cptr[j + 1]; ++p) {
            int32_t i = pp.crow[p];   // i > j, A(i, j) != 0
            if (first[j] <= maxfirst[i]) continue;   // j is not a leaf of the i-th row subtree
            maxfirst[i] = first[j];
            int32_t jprev = prevleaf[i];
            prevleaf[i] = j;
            if (jprev == -1) { ++delta[j]; continue; }   // first leaf
            int32_t q = jprev;
            while (q != anc[q]) q = anc[q];
            for (int32_t s = jprev; s != q;) { int32_t sp = anc[s]; anc[s] = q; s = sp; }
            ++delta[j];
            --delta[q];
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    // accumulate in an order where children precede parents
    for (int32_t k = 0; k < n; ++k) {
        int32_t j = post[k];
        if (parent[j] != -1) delta[parent[j]] += delta[j];
    }
    count.resize(n);
    for (int32_t j = 0; j < n; ++j) count[j] = static_cast<int32_t>(delta[j]);
}

}  // namespace

namespace {
struct StepClock {   // MADIPM_TIMING=1: wall time of every step of analyze on stderr (development aid)
    bool on = std::getenv("MADIPM_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[madipm analyze] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};
}  // namespace

void analyze(int32_t n, const int64_t* colptr, const int32_t* rowval, const SymbolicOptions& opt, Symbolic& S,
             const uint8_t* delayed) {
    S = Symbolic();
    S.n = n;
    S.nnz_lower = colptr[n];
    if (n == 0) { S.sn_first.assign(1, 0); S.sn_rowptr.assign(1, 0); S.sn_xoff.assign(1, 0); S.child_ptr.assign(1, 0); S.sn_coff.assign(1, 0); S.phase_ptr.assign(1, 0); S.big_ptr.assign(1, 0); return; }

    StepClock clk;
    // ---- 1. fill-reducing permutation
    std::vector<int32_t> perm0;
    if (opt.ordering == 1) { perm0.resize(n); std::iota(perm0.begin(), perm0.end(), 0); }
    else if (opt.ordering == 2) {
        // small leaves: the height of the supernodal tree (= number of phases of the factor / solve kernels, each a
        // latency-bound critical path) matters more than the few percent of fill that AMD on larger leaves saves
        int32_t leaf = opt.nd_leaf > 0 ? opt.nd_leaf : std::min(1024, std::max(128, n / 4096));
        nd_order(n, colptr, rowval, perm0, leaf, opt.nd_max_sep);
    } else amd_order(n, colptr, rowval, perm0, opt.dense_factor);
    std::vector<int32_t> ip0(n, -1);
    for (int32_t k = 0; k < n; ++k) {
        if (perm0[k] < 0 || perm0[k] >= n || ip0[perm0[k]] != -1) throw std::runtime_error("analyze: invalid permutation");
        ip0[perm0[k]] = k;
    }
    clk.lap("ordering");
    // ---- 1b. delay flagged nodes (K2 constraint rows) until just after a partner neighbour.
    // Partners are DISTINCT (a bipartite matching): at every stage the eliminated constraints then
    // have a structurally nonsingular block against the eliminated variables, so no pivot is
    // structurally zero even with delta_c = 0, and no pivot is the bare -1e-8 regularisation.
    if (delayed) {
        std::vector<int64_t> dptr(static_cast<size_t>(n) + 1, 0);
        for (int32_t j = 0; j < n; ++j)
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                int32_t i = rowval[p];
                if (i == j) continue;
                if (delayed[i] && !delayed[j]) ++dptr[i + 1];
                if (delayed[j] && !delayed[i]) ++dptr[j + 1];
            }
        for (int32_t v = 0; v < n; ++v) dptr[v + 1] += dptr[v];
        std::vector<int32_t> dadj(dptr[n]);
        {
            std::vector<int64_t> at(dptr.begin(), dptr.end() - 1);
            for (int32_t j = 0; j < n; ++j)
                for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                    int32_t i = rowval[p];
                    if (i == j) continue;
                    if (delayed[i] && !delayed[j]) dadj[at[i]++] = j;
                    if (delayed[j] && !delayed[i]) dadj[at[j]++] = i;
                }
        }
        std::vector<int32_t> dnodes;
        for (int32_t k = 0; k < n; ++k) if (delayed[perm0[k]]) dnodes.push_back(perm0[k]);   // in elimination order
        std::vector<int32_t> owner(n, -1), partner(n, -1), seen(n, -1), via(n, -1), queue;
        for (int32_t v : dnodes) {
            int32_t best = -1;
            for (int64_t p = dptr[v]; p < dptr[v + 1]; ++p) {
                int32_t x = dadj[p];
                if (owner[x] < 0 && (best < 0 || ip0[x] < ip0[best])) best = x;
            }
            if (best >= 0) { owner[best] = v; partner[v] = best; continue; }
            // augmenting path (breadth first over alternating paths)
            queue.clear();
            queue.push_back(v);
            int32_t found = -1;
            for (size_t qh = 0; qh < queue.size() && found < 0; ++qh) {
                int32_t u = queue[qh];
                for (int64_t p = dptr[u]; p < dptr[u + 1]; ++p) {
                    int32_t x = dadj[p];
                    if (seen[x] == v) continue;
                    seen[x] = v;
                    via[x] = u;
                    if (owner[x] < 0) { found = x; break; }
                    queue.push_back(owner[x]);
                }
            }
            while (found >= 0) {
                int32_t u = via[found];
                int32_t prev = partner[u];
                owner[found] = u;
                partner[u] = found;
                found = (u == v) ? -1 : prev;
            }
        }
        std::vector<int64_t> key(n);
        for (int32_t v = 0; v < n; ++v) {
            int64_t k = 2 * static_cast<int64_t>(ip0[v]);
            if (delayed[v] && partner[v] >= 0 && ip0[v] < ip0[partner[v]]) k = 2 * static_cast<int64_t>(ip0[partner[v]]) + 1;
            key[v] = k;
        }
        std::vector<int32_t> order(n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            return key[a] < key[b] || (key[a] == key[b] && ip0[a] < ip0[b]);
        });
        perm0 = order;
        for (int32_t k = 0; k < n; ++k) ip0[perm0[k]] = k;
    }
    clk.lap("constraint delay");
    // ---- 2. etree + counts in the AMD numbering, then one postorder that puts the
    //         heaviest child last (so it can chain into its parent's supernode)
    {
        PermutedPattern pp;
        permute_pattern(n, colptr, rowval, ip0, pp);
        std::vector<int32_t> par0, post0, cnt0, post1;
        etree(n, pp, par0);
        postorder(n, par0, nullptr, post0);
        column_counts(n, pp, par0, post0, cnt0);
        postorder(n, par0, &cnt0, post1);
        std::vector<int32_t> relabel(n);
        for (int32_t k = 0; k < n; ++k) relabel[post1[k]] = k;
        S.perm.resize(n);
        S.iperm.resize(n);
        S.parent.resize(n);
        S.colcount.resize(n);
        for (int32_t k = 0; k < n; ++k) {
            int32_t v = post1[k];
            S.perm[k] = perm0[v];
            S.parent[k] = par0[v] == -1 ? -1 : relabel[par0[v]];
            S.colcount[k] = cnt0[v];
        }
        for (int32_t k = 0; k < n; ++k) S.iperm[S.perm[k]] = k;
    }
    S.nnz_L = 0;
    S.flops = 0;
    for (int32_t j = 0; j < n; ++j) {
        S.nnz_L += S.colcount[j];
        S.flops += static_cast<double>(S.colcount[j]) * S.colcount[j];
        if (S.parent[j] != -1 && S.parent[j] <= j) throw std::runtime_error("analyze: etree not topologically numbered");
    }
    PermutedPattern pp;
    permute_pattern(n, colptr, rowval, S.iperm, pp);

    clk.lap("etree + counts + postorder");
    // ---- 3. supernodes: maximal chains parent[j] == j+1 with nested structure, width-capped
    const int32_t maxw = std::max(1, opt.max_width);
    std::vector<int32_t> first;   // first column of each initial supernode
    for (int32_t j = 0; j < n; ++j) {
        bool chain = j > 0 && S.parent[j - 1] == j && S.colcount[j - 1] == S.colcount[j] + 1 &&
                     (j - first.back()) < maxw;
        if (!chain) first.push_back(j);
    }
    int32_t ns0 = static_cast<int32_t>(first.size());
    first.push_back(n);
    std::vector<int32_t> col_sn(n);
    for (int32_t s = 0; s < ns0; ++s)
        for (int32_t j = first[s]; j < first[s + 1]; ++j) col_sn[j] = s;

    // ---- 4. relaxed amalgamation of the last child into its parent
    std::vector<int32_t> w(ns0), h(ns0), prev(ns0), owner(ns0), lastcol(ns0);
    std::vector<int64_t> nzt(ns0);
    std::vector<char> dead(ns0, 0);
    for (int32_t s = 0; s < ns0; ++s) {
        w[s] = first[s + 1] - first[s];
        h[s] = S.colcount[first[s]];
        lastcol[s] = first[s + 1] - 1;
        nzt[s] = 0;
        for (int32_t j = first[s]; j < first[s + 1]; ++j) nzt[s] += S.colcount[j];
        prev[s] = s - 1;
        owner[s] = s;
    }
    auto find = [&](int32_t s) {
        while (owner[s] != s) { owner[s] = owner[owner[s]]; s = owner[s]; }
        return s;
    };
    for (int32_t p = 0; p < ns0; ++p) {
        for (;;) {
            int32_t q = prev[p];
            if (q < 0) break;
            int32_t pc = S.parent[lastcol[q]];
            if (pc == -1 || find(col_sn[pc]) != p) break;
            int32_t wm = w[q] + w[p];
            if (wm > maxw) break;
            int32_t hm = w[q] + h[p];
            double stored = static_cast<double>(wm) * hm - 0.5 * wm * (wm - 1.0);
            double zeros = stored - static_cast<double>(nzt[q] + nzt[p]);
            double z = zeros / stored;
            bool merge;
            if (wm <= opt.relax_small) merge = true;
            else if (wm <= 16) merge = z <= opt.relax_zeros[0];
            else if (wm <= 48) merge = z <= opt.relax_zeros[1];
            else merge = z <= opt.relax_zeros[2];
            if (!merge) break;
            // q is absorbed by p
            dead[q] = 1;
            owner[q] = p;
            first[p] = first[q];
            w[p] = wm;
            h[p] = hm;
            nzt[p] += nzt[q];
            prev[p] = prev[q];
        }
    }
    S.sn_first.clear();
    for (int32_t s = 0; s < ns0; ++s)
        if (!dead[s]) S.sn_first.push_back(first[s]);
    S.nsuper = static_cast<int32_t>(S.sn_first.size());
    S.sn_first.push_back(n);
    const int32_t ns = S.nsuper;
    S.sn_of.resize(n);
    for (int32_t s = 0; s < ns; ++s)
        for (int32_t j = S.sn_first[s]; j < S.sn_first[s + 1]; ++j) S.sn_of[j] = s;
    S.sn_parent.assign(ns, -1);
    for (int32_t s = 0; s < ns; ++s) {
        int32_t pc = S.parent[S.sn_first[s + 1] - 1];
        S.sn_parent[s] = pc == -1 ? -1 : S.sn_of[pc];
    }

    clk.lap("supernodes + amalgamation");
    // ---- 5. row structures: own columns, then the union of A's columns and the children's tails
    std::vector<int32_t> chead(ns, -1), cnext(ns, -1);
    for (int32_t s = ns - 1; s >= 0; --s)
        if (S.sn_parent[s] != -1) { cnext[s] = chead[S.sn_parent[s]]; chead[S.sn_parent[s]] = s; }
    S.sn_rowptr.assign(static_cast<size_t>(ns) + 1, 0);
    S.sn_rows.clear();
    S.sn_rows.reserve(static_cast<size_t>(S.nnz_L / 2 + n));
    std::vector<int32_t> mark(n, -1);
    std::vector<int32_t> tail;
    for (int32_t s = 0; s < ns; ++s) {
        const int32_t fc = S.sn_first[s], lc = S.sn_first[s + 1] - 1;
        tail.clear();
        for (int32_t j = fc; j <= lc; ++j)
            for (int64_t p = pp.cptr[j]; p < pp.cptr[j + 1]; ++p) {
                int32_t r = pp.crow[p];
                if (r > lc && mark[r] != s) { mark[r] = s; tail.push_back(r); }
            }
        for (int32_t c = chead[s]; c != -1; c = cnext[c]) {
            const int32_t wc = S.sn_first[c + 1] - S.sn_first[c];
            for (int64_t p = S.sn_rowptr[c] + wc; p < S.sn_rowptr[c + 1]; ++p) {
                int32_t r = S.sn_rows[p];
                if (r > lc && mark[r] != s) { mark[r] = s; tail.push_back(r); }
            }
        }
        std::sort(tail.begin(), tail.end());
        for (int32_t j = fc; j <= lc; ++j) S.sn_rows.push_back(j);
        S.sn_rows.insert(S.sn_rows.end(), tail.begin(), tail.end());
        S.sn_rowptr[s + 1] = static_cast<int64_t>(S.sn_rows.size());
    }

    clk.lap("row structures");
    // ---- 6. panel offsets (h x w column-major, start aligned to 16 bytes)
    S.sn_xoff.assign(static_cast<size_t>(ns) + 1, 0);
    S.max_h = S.max_w = 0;
    for (int32_t s = 0; s < ns; ++s) {
        int64_t hs = S.sn_rowptr[s + 1] - S.sn_rowptr[s];
        int64_t ws = S.sn_first[s + 1] - S.sn_first[s];
        S.max_h = std::max<int32_t>(S.max_h, static_cast<int32_t>(hs));
        S.max_w = std::max<int32_t>(S.max_w, static_cast<int32_t>(ws));
        int64_t sz = hs * ws;
        S.sn_xoff[s + 1] = S.sn_xoff[s] + ((sz + 1) & ~int64_t(1));
    }
    S.nnz_stored = S.sn_xoff[ns];

    clk.lap("panel offsets");
    // ---- 7. assembly map: CSC nz (i, j) -> slot of (max(pi,pj), min(pi,pj)) in its panel.
    // Independent per column: computed by a few host threads while steps 7b and 8 run on this one (joined before return).
    S.asm_map.assign(S.nnz_lower, -1);
    std::exception_ptr asm_err, sc_err;       // (declared before the joiner below: the threads write them)
    std::atomic<bool> asm_failed{false};
    auto asm_range = [&](int32_t j0, int32_t j1) {
        try {
            for (int32_t j = j0; j < j1; ++j)
                for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                    int32_t a = S.iperm[rowval[p]], b = S.iperm[j];
                    int32_t r = std::max(a, b), c = std::min(a, b);
                    int32_t s = S.sn_of[c];
                    const int32_t* rb = S.sn_rows.data() + S.sn_rowptr[s];
                    const int32_t* re = S.sn_rows.data() + S.sn_rowptr[s + 1];
                    const int32_t* it = std::lower_bound(rb, re, r);
                    if (it == re || *it != r) throw std::runtime_error("analyze: entry outside the symbolic structure");
                    int64_t hs = re - rb;
                    S.asm_map[p] = S.sn_xoff[s] + static_cast<int64_t>(c - S.sn_first[s]) * hs + (it - rb);
                }
        } catch (...) {
            if (!asm_failed.exchange(true)) asm_err = std::current_exception();
        }
    };
    std::vector<std::thread> asm_threads;
    {
        const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
        const int32_t nt = n < 50000 ? 0 : static_cast<int32_t>(std::min(4u, hc > 1 ? hc - 1 : 1u));
        if (nt == 0) asm_range(0, n);
        for (int32_t t = 0; t < nt; ++t)
            asm_threads.emplace_back(asm_range, static_cast<int32_t>(static_cast<int64_t>(n) * t / nt),
                                     static_cast<int32_t>(static_cast<int64_t>(n) * (t + 1) / nt));
    }
    struct Joiner {   // the threads are joined on every way out of analyze
        std::vector<std::thread>& ts;
        ~Joiner() { for (auto& t : ts) if (t.joinable()) t.join(); }
    } asm_join{asm_threads};
    if (asm_failed.load() && asm_threads.empty()) std::rethrow_exception(asm_err);

    clk.lap("assembly map");
    // ---- 7b. scatter tables for the right-looking update (reads the row structures only; step 8 does not read the
    // tables, so for large patterns they are filled by one more host thread while the schedule is built)
    auto scatter_tables = [&]() {
      try {
        const size_t nr = S.sn_rows.size();
        S.row_colbase.assign(nr, 0);
        S.row_relbase.assign(nr, Symbolic::kRelSearch);
        S.rel.assign(4096, 0);   // padding keeps every table base (size - jb, jb <= 4096) non-negative
        // Tall supernodes (more than 4096 rows) get tabulated positions too while the tables stay below a budget
        // (default 1e9 entries = 4 GB of the 180 GB): the alternative is a binary search per updated entry.
        const char* benv = std::getenv("MADIPM_REL_BUDGET");
        const double rel_budget = benv ? std::atof(benv) : 1.0e9;
        const char* cenv = std::getenv("MADIPM_CHECK_TABLES");
        const bool check_tables = cenv && cenv[0] == '1';
        std::vector<int32_t> pos;
        for (int32_t s = 0; s < ns; ++s) {
            const int64_t r0 = S.sn_rowptr[s];
            const int32_t hs = static_cast<int32_t>(S.sn_rowptr[s + 1] - r0);
            const int32_t ws = S.sn_first[s + 1] - S.sn_first[s];
            const int32_t* rows = S.sn_rows.data() + r0;
            int32_t jb = ws;
            int32_t prev_t = -1, prev_ht = 0, prev_w = 0, prev_jb = 0;
            bool prev_full = false;       // `pos` holds the positions in prev_t of rows[prev_jb .. hs)
            while (jb < hs) {
                const int32_t t = S.sn_of[rows[jb]];
                const int32_t fct = S.sn_first[t], lct = S.sn_first[t + 1];
                const int32_t* trows = S.sn_rows.data() + S.sn_rowptr[t];
                const int32_t ht = static_cast<int32_t>(S.sn_rowptr[t + 1] - S.sn_rowptr[t]);
                int32_t je = jb;
                while (je < hs && rows[je] < lct) ++je;
                // positions of rows[jb..hs) in trows (both ascending).  Fast path: the rows are a subset of the ancestor's
                // (a property of the symbolic factorisation), so when the ancestor has exactly as many rows from the position
                // of rows[jb] on, the two lists coincide there and every position is i + c -- the whole dense trailing chain
                // of a front-dominated problem (C4: 10^10 binary-search steps otherwise).
                bool shift = true;
                int32_t pos0 = 0;
                {
                    const int32_t* it0 = std::lower_bound(trows, trows + ht, rows[jb]);
                    if (it0 == trows + ht || *it0 != rows[jb]) throw std::runtime_error("analyze: update row missing in ancestor");
                    pos0 = static_cast<int32_t>(it0 - trows);
                }
                if (ht - pos0 != hs - jb || trows[ht - 1] != rows[hs - 1]) {
                    if (prev_t != -1 && S.sn_parent[prev_t] == t && ht == prev_ht - prev_w && prev_full) {
                        // t's row list is exactly the below-part of the previous target's (parent with as many rows as the
                        // child has below its columns): positions shift by the previous target's width
                        const int32_t skip = jb - prev_jb;
                        for (int32_t i = jb; i < hs; ++i) pos[i - jb] = pos[i - jb + skip] - prev_w;
                        pos.resize(hs - jb);
                        for (int32_t i = jb; i < hs && shift; ++i)
                            if (pos[i - jb] - pos[0] != i - jb) shift = false;
                    } else {
                        pos.resize(hs - jb);
                        const int32_t* from = trows;
                        for (int32_t i = jb; i < hs; ++i) {
                            const int32_t* it = std::lower_bound(from, trows + ht, rows[i]);
                            if (it == trows + ht || *it != rows[i]) throw std::runtime_error("analyze: update row missing in ancestor");
                            pos[i - jb] = static_cast<int32_t>(it - trows);
                            if (pos[i - jb] - pos[0] != i - jb) shift = false;
                            from = it + 1;
                        }
                    }
                    prev_full = true;
                } else {
                    prev_full = false;
                }
                prev_t = t; prev_ht = ht; prev_w = lct - fct; prev_jb = jb;
                if (check_tables) {                       // MADIPM_CHECK_TABLES=1: every shortcut above against the plain search
                    for (int32_t i = jb; i < hs; ++i) {
                        const int32_t* it = std::lower_bound(trows, trows + ht, rows[i]);
                        const int32_t want = static_cast<int32_t>(it - trows);
                        const int32_t got = prev_full ? pos[i - jb] : pos0 + (i - jb);
                        if (it == trows + ht || *it != rows[i] || got != want || (shift && want - pos0 != i - jb))
                            throw std::runtime_error("analyze: scatter-table shortcut disagrees with the search");
                    }
                }
                int64_t relbase;
                if (shift) {
                    const int64_t c = static_cast<int64_t>(pos0) - jb;       // position = i + c
                    relbase = -(c + Symbolic::kRelShift) - 1;
                } else if (hs <= 4096 || static_cast<double>(S.rel.size()) + (hs - jb) <= rel_budget) {
                    relbase = static_cast<int64_t>(S.rel.size()) - jb;      // rel[relbase + i]
                    S.rel.insert(S.rel.end(), pos.begin(), pos.end());
                } else {
                    relbase = Symbolic::kRelSearch;
                }
                for (int32_t jj = jb; jj < je; ++jj) {
                    S.row_colbase[r0 + jj] = S.sn_xoff[t] + static_cast<int64_t>(rows[jj] - fct) * ht;
                    S.row_relbase[r0 + jj] = relbase;
                }
                jb = je;
            }
        }
    
      } catch (...) { sc_err = std::current_exception(); }
    };
    if (n >= 50000) asm_threads.emplace_back(scatter_tables);
    else { scatter_tables(); if (sc_err) std::rethrow_exception(sc_err); }

    clk.lap("scatter tables");
    // ---- 8. schedule by height (see symbolic.hpp)
    S.sn_level.assign(ns, 0);
    for (int32_t s = 0; s < ns; ++s) {
        int32_t p = S.sn_parent[s];
        if (p != -1) {
            if (p <= s) throw std::runtime_error("analyze: supernodal etree not topologically numbered");
            S.sn_level[p] = std::max(S.sn_level[p], S.sn_level[s] + 1);
        }
    }
    S.n_levels = 0;
    for (int32_t s = 0; s < ns; ++s) S.n_levels = std::max(S.n_levels, S.sn_level[s] + 1);
    S.child_ptr.assign(static_cast<size_t>(ns) + 1, 0);
    for (int32_t s = 0; s < ns; ++s)
        if (S.sn_parent[s] != -1) ++S.child_ptr[S.sn_parent[s] + 1];
    for (int32_t s = 0; s < ns; ++s) S.child_ptr[s + 1] += S.child_ptr[s];
    S.child_idx.assign(S.child_ptr[ns], 0);
    {
        std::vector<int32_t> at(S.child_ptr.begin(), S.child_ptr.end() - 1);
        for (int32_t s = 0; s < ns; ++s)
            if (S.sn_parent[s] != -1) S.child_idx[at[S.sn_parent[s]]++] = s;
    }
    auto is_big = [&](int32_t s) {
        int64_t hs = S.sn_rowptr[s + 1] - S.sn_rowptr[s], ws = S.sn_first[s + 1] - S.sn_first[s];
        return hs * ws > opt.block_panel || hs - ws > opt.block_below;
    };
    S.sn_kind.assign(ns, 0);
    for (int32_t s = 0; s < ns; ++s) {
        if (is_big(s)) S.sn_kind[s] = 2;
        else if (S.sn_parent[s] != -1 && is_big(S.sn_parent[s])) S.sn_kind[s] = 1;
    }
    // position of every below-row inside the parent's row list (rows of a child below its own columns are a
    // subset of the parent's rows; both lists ascend)
    S.par_pos.assign(S.sn_rows.size(), 0);
    for (int32_t s = 0; s < ns; ++s) {
        const int32_t p = S.sn_parent[s];
        if (p == -1) continue;
        const int32_t ws = S.sn_first[s + 1] - S.sn_first[s];
        const int32_t* prow = S.sn_rows.data() + S.sn_rowptr[p];
        const int32_t hp = static_cast<int32_t>(S.sn_rowptr[p + 1] - S.sn_rowptr[p]);
        int32_t q = 0;
        for (int64_t k = S.sn_rowptr[s] + ws; k < S.sn_rowptr[s + 1]; ++k) {
            while (q < hp && prow[q] < S.sn_rows[k]) ++q;
            if (q == hp || prow[q] != S.sn_rows[k]) throw std::runtime_error("analyze: child row missing in parent");
            S.par_pos[k] = q;
        }
    }
    S.sn_coff.assign(static_cast<size_t>(ns) + 1, 0);
    S.small_max_panel = S.small_max_h = 0;
    for (int32_t s = 0; s < ns; ++s) {
        int64_t hs = S.sn_rowptr[s + 1] - S.sn_rowptr[s], ws = S.sn_first[s + 1] - S.sn_first[s];
        int64_t nb = hs - ws, sz = 0;
        if (S.sn_kind[s] != 2) {
            sz = (nb * (nb + 1) / 2 + 1) & ~int64_t(1);
            S.small_max_panel = std::max<int32_t>(S.small_max_panel, static_cast<int32_t>(hs * ws));
            S.small_max_h = std::max<int32_t>(S.small_max_h, static_cast<int32_t>(hs));
        }
        S.sn_coff[s + 1] = S.sn_coff[s] + sz;
    }
    // shared-memory slices of the four task classes (doubles): [panel | h doubles | max_w doubles | 16 child records | ticket]
    auto even = [](int64_t v) { return static_cast<int32_t>((v + 1) & ~int64_t(1)); };
    auto panel_need = [&](int64_t np) { return even(np) + 2; };            // odd tail chunk + 4-column tile overrun
    S.tree_fixed = even(std::max(S.max_w, 1)) + 16 * 4 + 2;
    // a multiple of 16 doubles, so that the slices of the smaller classes nest exactly inside the larger ones
    // (the h-doubles region doubles as the column buffer of dense_ldl_blk: at least 24 doubles)
    const int32_t slice256 = (panel_need(std::max(S.small_max_panel, 2)) + even(std::max(S.small_max_h, 24)) + S.tree_fixed + 15) & ~15;
    for (int32_t c = 0; c < 4; ++c) {
        const int32_t G = 256 >> c;
        S.tree_slice[c] = slice256 >> c;
        // a group of G threads owns at most two rows per thread in the panel step of dense_ldl_blk
        S.tree_maxh[c] = c == 0 ? even(std::max(S.small_max_h, 24)) : std::min(even(std::max(S.small_max_h, 24)), 2 * G);
        S.tree_cap[c] = S.tree_slice[c] - S.tree_maxh[c] - S.tree_fixed;
    }
    std::vector<int8_t> cls(ns, 0);
    for (int32_t s = 0; s < ns; ++s) {
        if (S.sn_kind[s] == 2) continue;
        int64_t hs = S.sn_rowptr[s + 1] - S.sn_rowptr[s], ws = S.sn_first[s + 1] - S.sn_first[s];
        for (int32_t c = 3; c >= 1; --c)
            if (hs <= S.tree_maxh[c] && panel_need(hs * ws) <= S.tree_cap[c]) { cls[s] = static_cast<int8_t>(c); break; }
    }
    S.phase_ptr.assign(static_cast<size_t>(S.n_levels) + 1, 0);
    S.big_ptr.assign(static_cast<size_t>(S.n_levels) + 1, 0);
    S.phase_nblock.assign(S.n_levels, 0);
    S.phase_cls.assign(static_cast<size_t>(S.n_levels) * 5, 0);
    for (int32_t s = 0; s < ns; ++s) {
        if (S.sn_kind[s] == 2) ++S.big_ptr[S.sn_level[s] + 1];
        else { ++S.phase_ptr[S.sn_level[s] + 1]; if (cls[s] != 3) ++S.phase_nblock[S.sn_level[s]]; }
    }
    for (int32_t l = 0; l < S.n_levels; ++l) { S.phase_ptr[l + 1] += S.phase_ptr[l]; S.big_ptr[l + 1] += S.big_ptr[l]; }
    S.task_sn.assign(S.phase_ptr[S.n_levels], 0);
    S.big_sn.assign(S.big_ptr[S.n_levels], 0);
    {
        std::vector<int32_t> at(S.phase_ptr.begin(), S.phase_ptr.end() - 1), bat(S.big_ptr.begin(), S.big_ptr.end() - 1);
        for (int32_t s = 0; s < ns; ++s) {
            if (S.sn_kind[s] == 2) S.big_sn[bat[S.sn_level[s]]++] = s;
            else S.task_sn[at[S.sn_level[s]]++] = s;
        }
        auto cost = [&](int32_t s) {
            double hs = static_cast<double>(S.sn_rowptr[s + 1] - S.sn_rowptr[s]), ws = static_cast<double>(S.sn_first[s + 1] - S.sn_first[s]);
            return ws * hs * hs;
        };
        for (int32_t l = 0; l < S.n_levels; ++l) {
            std::stable_sort(S.task_sn.begin() + S.phase_ptr[l], S.task_sn.begin() + S.phase_ptr[l + 1], [&](int32_t a, int32_t b) {
                if (cls[a] != cls[b]) return cls[a] < cls[b];
                return cost(a) > cost(b);
            });
            int32_t q = S.phase_ptr[l];
            for (int32_t c = 0; c < 4; ++c) {
                S.phase_cls[static_cast<size_t>(l) * 5 + c] = q;
                while (q < S.phase_ptr[l + 1] && cls[S.task_sn[q]] == c) ++q;
            }
            S.phase_cls[static_cast<size_t>(l) * 5 + 4] = S.phase_ptr[l + 1];
        }
    }
    S.n_big = static_cast<int32_t>(S.big_sn.size());
    S.n_block_tasks = 0;
    for (int32_t l = 0; l < S.n_levels; ++l) S.n_block_tasks += S.phase_nblock[l];
    S.n_warp_tasks = static_cast<int32_t>(S.task_sn.size()) - S.n_block_tasks;
    clk.lap("schedule");
    for (auto& t : asm_threads) t.join();
    if (asm_failed.load()) std::rethrow_exception(asm_err);
    if (sc_err) std::rethrow_exception(sc_err);
    clk.lap("maps + tables (joined)");
}

}  // namespace madipm
