// Nested dissection by level-structure separators with AMD on the leaves.
//
// Why it exists: minimum-degree orderings of banded / path-like KKT graphs give a
// caterpillar elimination tree (thousands of sequential levels), which a GPU can
// only walk one small front at a time.  Dissection gives a balanced tree: the
// leaves become independent bottom subtrees (one thread block each) and only
// O(log n) separator levels remain.  The separators are the middle level of a
// rooted level structure from a pseudo-peripheral node (George & Liu, "Computer
// Solution of Large Sparse Positive Definite Systems", 1981, ch. 8), trimmed to
// the vertices that touch the next level.  When a region has no thin separator
// (expander-like graphs) it is handed to AMD whole, so the hybrid never does
// worse than AMD by more than the separators already placed.
//
// The two halves of a dissection are independent, so the first few levels of the recursion run them on separate host
// threads.  The result does not depend on the schedule: a region of k vertices owns a fixed range of k positions of the
// order (first half, second half, separator), region labels and BFS stamps come from atomic counters (only their
// uniqueness matters), and all per-vertex scratch is touched only for the vertices of the task's own region.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <exception>
#include <numeric>
#include <stdexcept>
#include <thread>
#include <vector>

#include "symbolic.hpp"

namespace madipm {

namespace {

struct Graph {
    int32_t n;
    std::vector<int64_t> ptr;
    std::vector<int32_t> adj;
};

struct Dissector {
    const Graph& g;
    int32_t leaf_size;
    double max_sep_frac;
    int par_depth;                                // recursion levels that fork a host thread
    std::vector<std::atomic<int32_t>> region;     // region id of every vertex (-1 = already ordered); neighbours in other
                                                  // regions are read while their owner relabels them (never to our id)
    std::vector<int32_t> level;    // BFS level scratch (valid only for stamped vertices)
    std::vector<int32_t> stamp;
    std::atomic<int32_t> stamp_counter{0};
    std::atomic<int32_t> next_region{1};
    std::vector<int32_t> order;    // output: elimination order (old indices), written by position
    std::vector<int32_t> local;    // old -> local index scratch for leaf AMD

    Dissector(const Graph& gr, int32_t leaf, double frac, int pdepth)
        : g(gr), leaf_size(leaf), max_sep_frac(frac), par_depth(pdepth), region(gr.n), level(gr.n, 0), stamp(gr.n, 0),
          order(gr.n, -1), local(gr.n, -1) {
        for (auto& r : region) r.store(0, std::memory_order_relaxed);
    }

    int32_t reg(int32_t v) const { return region[v].load(std::memory_order_relaxed); }
    void set_reg(int32_t v, int32_t r) { region[v].store(r, std::memory_order_relaxed); }

    // BFS inside region `rid` from `root`; fills `out` in BFS order, returns #levels; `st` = the stamp it used
    int32_t bfs(int32_t root, int32_t rid, std::vector<int32_t>& out, std::vector<int32_t>& level_start, int32_t& st) {
        st = stamp_counter.fetch_add(1, std::memory_order_relaxed) + 1;
        out.clear();
        level_start.clear();
        out.push_back(root);
        stamp[root] = st;
        level[root] = 0;
        size_t head = 0;
        level_start.push_back(0);
        int32_t cur = 0;
        while (head < out.size()) {
            int32_t v = out[head];
            if (level[v] != cur) { cur = level[v]; level_start.push_back(static_cast<int32_t>(head)); }
            ++head;
            for (int64_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
                int32_t u = g.adj[p];
                if (reg(u) == rid && stamp[u] != st) {
                    stamp[u] = st;
                    level[u] = level[v] + 1;
                    out.push_back(u);
                }
            }
        }
        level_start.push_back(static_cast<int32_t>(out.size()));
        return static_cast<int32_t>(level_start.size()) - 1;
    }

    // the vertices of a leaf take positions [base, base + |verts|) of the order
    void order_leaf(const std::vector<int32_t>& verts, int64_t base) {
        const int32_t k = static_cast<int32_t>(verts.size());
        if (k <= 2) {
            for (int32_t i = 0; i < k; ++i) { order[base + i] = verts[i]; set_reg(verts[i], -1); }
            return;
        }
        for (int32_t i = 0; i < k; ++i) local[verts[i]] = i;
        std::vector<int64_t> cp(static_cast<size_t>(k) + 1, 0);
        std::vector<int32_t> ri;
        for (int32_t i = 0; i < k; ++i) {
            int32_t v = verts[i];
            for (int64_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
                int32_t u = g.adj[p];
                if (local[u] > i) ri.push_back(local[u]);   // strict lower triangle is enough
            }
            cp[i + 1] = static_cast<int64_t>(ri.size());
        }
        std::vector<int32_t> p;
        amd_order(k, cp.data(), ri.data(), p, 10.0);
        for (int32_t i = 0; i < k; ++i) order[base + i] = verts[p[i]];
        for (int32_t v : verts) { local[v] = -1; set_reg(v, -1); }
    }

    void dissect(std::vector<int32_t> verts, int32_t rid, int depth, int64_t base) {
        const int32_t k = static_cast<int32_t>(verts.size());
        if (k == 0) return;
        if (k <= leaf_size || depth > 48) { order_leaf(verts, base); return; }
        // connected components of the region
        std::vector<int32_t> bfs_out, lstart;
        int32_t st = 0;
        bfs(verts[0], rid, bfs_out, lstart, st);
        if (static_cast<int32_t>(bfs_out.size()) < k) {
            std::vector<std::pair<std::vector<int32_t>, int32_t>> comps;
            for (int32_t v0 : verts) {
                if (reg(v0) != rid) continue;
                std::vector<int32_t> comp, ls;
                int32_t st2 = 0;
                bfs(v0, rid, comp, ls, st2);
                int32_t crid = next_region.fetch_add(1, std::memory_order_relaxed);
                for (int32_t v : comp) set_reg(v, crid);
                comps.emplace_back(std::move(comp), crid);
            }
            verts.clear();
            verts.shrink_to_fit();
            // biggest last, so the heavy component sits next to whatever follows
            std::stable_sort(comps.begin(), comps.end(),
                             [](const auto& a, const auto& b) { return a.first.size() < b.first.size(); });
            int64_t at = base;
            for (auto& c : comps) {
                const int64_t sz = static_cast<int64_t>(c.first.size());
                dissect(std::move(c.first), c.second, depth + 1, at);
                at += sz;
            }
            return;
        }
        // pseudo-peripheral root: walk to a min-degree vertex of the last level while depth grows
        int32_t nlev = static_cast<int32_t>(lstart.size()) - 1;
        for (int it = 0; it < 6; ++it) {
            int32_t best = -1;
            int64_t bestdeg = INT64_MAX;
            for (int32_t q = lstart[nlev - 1]; q < lstart[nlev]; ++q) {
                int32_t v = bfs_out[q];
                int64_t d = g.ptr[v + 1] - g.ptr[v];
                if (d < bestdeg) { bestdeg = d; best = v; }
            }
            std::vector<int32_t> out2, ls2;
            int32_t st2 = 0;
            int32_t nl2 = bfs(best, rid, out2, ls2, st2);
            if (nl2 > nlev) { nlev = nl2; bfs_out.swap(out2); lstart.swap(ls2); st = st2; }
            else { // keep the deeper structure; level[] and stamp[] must match bfs_out again: the trial BFS overwrote
                   // them, and (bfs_out, lstart) hold everything a BFS from `root` would write, so one linear pass
                   // restores them (same values as the BFS it replaces)
                st = stamp_counter.fetch_add(1, std::memory_order_relaxed) + 1;
                for (int32_t l = 0; l < nlev; ++l)
                    for (int32_t q = lstart[l]; q < lstart[l + 1]; ++q) { level[bfs_out[q]] = l; stamp[bfs_out[q]] = st; }
                break;
            }
        }
        if (nlev < 3) { order_leaf(verts, base); return; }
        // middle level by vertex count
        int32_t mid = 1;
        {
            int64_t best = INT64_MAX;
            for (int32_t j = 1; j <= nlev - 2; ++j) {
                int64_t below = lstart[j];
                int64_t above = k - lstart[j + 1];
                int64_t imb = below > above ? below - above : above - below;
                if (imb < best) { best = imb; mid = j; }
            }
        }
        // separator: vertices of level `mid` adjacent to level mid + 1
        std::vector<int32_t> sep, partA, partB;
        for (int32_t q = lstart[mid]; q < lstart[mid + 1]; ++q) {
            int32_t v = bfs_out[q];
            bool touches = false;
            for (int64_t p = g.ptr[v]; p < g.ptr[v + 1] && !touches; ++p) {
                int32_t u = g.adj[p];
                if (reg(u) == rid && stamp[u] == st && level[u] == mid + 1) touches = true;
            }
            if (touches) sep.push_back(v);
            else partA.push_back(v);
        }
        if (static_cast<double>(sep.size()) > max_sep_frac * k) { order_leaf(verts, base); return; }
        for (int32_t q = 0; q < lstart[mid]; ++q) partA.push_back(bfs_out[q]);
        for (int32_t q = lstart[mid + 1]; q < k; ++q) partB.push_back(bfs_out[q]);
        int32_t ra = next_region.fetch_add(2, std::memory_order_relaxed), rb = ra + 1;
        for (int32_t v : partA) set_reg(v, ra);
        for (int32_t v : partB) set_reg(v, rb);
        for (int32_t v : sep) set_reg(v, -1);
        verts.clear();
        verts.shrink_to_fit();
        const int64_t na = static_cast<int64_t>(partA.size()), nb = static_cast<int64_t>(partB.size());
        if (depth < par_depth && std::min(na, nb) >= 4096) {
            std::exception_ptr err;
            std::thread ta([&, this]() {
                try { dissect(std::move(partA), ra, depth + 1, base); } catch (...) { err = std::current_exception(); }
            });
            try { dissect(std::move(partB), rb, depth + 1, base + na); } catch (...) { ta.join(); throw; }
            ta.join();
            if (err) std::rethrow_exception(err);
        } else {
            dissect(std::move(partA), ra, depth + 1, base);
            dissect(std::move(partB), rb, depth + 1, base + na);
        }
        for (size_t i = 0; i < sep.size(); ++i) order[base + na + nb + static_cast<int64_t>(i)] = sep[i];
    }
};

}  // namespace

void nd_order(int32_t n, const int64_t* colptr, const int32_t* rowval, std::vector<int32_t>& perm,
              int32_t leaf_size, double max_sep_frac) {
    perm.assign(n, 0);
    if (n == 0) return;
    Graph g;
    g.n = n;
    g.ptr.assign(static_cast<size_t>(n) + 1, 0);
    for (int32_t j = 0; j < n; ++j)
        for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
            int32_t i = rowval[p];
            if (i == j) continue;
            ++g.ptr[i + 1];
            ++g.ptr[j + 1];
        }
    for (int32_t i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
    g.adj.assign(g.ptr[n], 0);
    {
        std::vector<int64_t> at(g.ptr.begin(), g.ptr.end() - 1);
        for (int32_t j = 0; j < n; ++j)
            for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
                int32_t i = rowval[p];
                if (i == j) continue;
                g.adj[at[i]++] = j;
                g.adj[at[j]++] = i;
            }
    }
    // host threads for the first levels of the recursion: 2^depth leaves of work, at most 16 threads;
    // MADIPM_ND_THREADS=1 keeps everything on the calling thread (same ordering either way)
    int par_depth = 0;
    {
        unsigned hc = std::thread::hardware_concurrency();
        if (const char* e = std::getenv("MADIPM_ND_THREADS")) hc = static_cast<unsigned>(std::max(1, std::atoi(e)));
        while ((2u << par_depth) <= hc && par_depth < 4) ++par_depth;
    }
    Dissector d(g, std::max(8, leaf_size), max_sep_frac, par_depth);
    std::vector<int32_t> all(n);
    std::iota(all.begin(), all.end(), 0);
    d.dissect(std::move(all), 0, 0, 0);
    for (int32_t v : d.order)
        if (v < 0) throw std::runtime_error("nd_order: incomplete ordering");
    perm = d.order;
}

}  // namespace madipm
