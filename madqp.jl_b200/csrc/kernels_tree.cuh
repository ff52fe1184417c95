// Height-scheduled supernodal LDL' and triangular solves for the "small" supernodes (north-star part (b)).
//
// All supernodes of one height of the supernodal elimination tree are independent.  The three kernels below
// are persistent: one cooperative launch walks a range of heights, with a grid-wide barrier between two
// heights, so a factorisation or a triangular sweep of a tree whose panels all fit one block is ONE launch.
// Inside a height every supernode is a task for a GROUP of 256, 128, 64 or 32 threads (its class: the smallest
// group whose proportional slice of the block's shared memory holds the panel), staged in shared memory by
// cp.async (16-byte, L2 only).  Groups take tasks from a per-(height, class) ticket counter, largest class first,
// so a block runs one 256-thread task, or two 128-thread tasks, ... or eight warp tasks at a time and nobody idles
// while work of its class is left.  The last heights of a tree hold fewer tasks than the GPU has SMs; they run in
// a second kernel with one 1024-thread block per task (k_tree_top): there the latency of a single task is all that
// matters, and four times the threads put four times the loads in flight.
//
// The factorisation is multifrontal in its data flow, supernodal in its storage:
//   * a task loads its h x w panel (original KKT entries, put there by k_assemble_panels, plus whatever big
//     descendants pushed), adds the parts of its children's contribution blocks that land in the panel,
//     factorises it (dense LDL', D on the diagonal, unit-lower L), and stores it;
//   * it then writes its own contribution block C = sum(children's pass-through parts) - L21 D L21' (packed
//     lower triangle) for the parent to pull.  Nothing is shared between concurrent tasks: no atomics, and
//     the summation order is fixed, so the factor is bit-reproducible from run to run.
//   * only where a small supernode sits under a big (multi-block) parent does it scatter C into the
//     ancestors' panels itself (kind 1, FP64 atomics through the tables of kernels_factor.cuh).
// The forward sweep moves contribution VECTORS the same way (parent pulls), the backward sweep only reads.
//
// Everything another block may have written during the same launch (panels, C, cvec, xw) is read with
// ld.global.cg / cp.async.cg; index tables are read-only and go through the read-only path.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "kernels_factor.cuh"
#include "dense_ldl_blk.cuh"

namespace madipm {

constexpr int kTreeThreads = 256;
constexpr int kTreeWarps = kTreeThreads / 32;

constexpr int kChildChunk = 16;  // child records staged per pass

// Everything a task needs to know about its supernode, in one 48-byte record stored in TASK order (one memory round
// trip instead of a chain of dependent index loads: the tasks are latency-bound, not bandwidth-bound).
struct __align__(16) SnMeta {
    int64_t xoff;      // panel offset in Lx
    int64_t coff;      // contribution-block offset in Cx
    int32_t rowptr;    // offset of the row list in sn_rows / par_pos / cvec
    int32_t fc;        // first column
    int32_t h;         // rows of the panel
    int32_t w;         // columns of the panel
    int32_t c0, c1;    // its child records are crec[c0 .. c1)
    int32_t sn;        // supernode number
    int32_t kind;      // 0 small (parent pulls), 1 small under a big parent
};
// One (parent, child) edge, stored in child_ptr order.
struct __align__(16) ChildRec {
    int64_t coff;      // the child's contribution block
    int32_t below;     // offset of the child's first below-row in sn_rows / par_pos / cvec
    int32_t nb;        // number of below rows of the child
    int32_t nbp;       // ... of which fall inside the parent's own columns (they come first)
    int32_t pull;      // 1: the child left a contribution block / vector for the parent (kind 0)
    int32_t pad[2];
};

// Sync record of one height: tickets of the four task classes, then the grid-barrier counter of the launch that STARTS at
// this height (32 bytes, so one cudaMemsetAsync over the heights of a launch clears everything the launch synchronises on).
constexpr int kSyncStride = 8;
constexpr int kSyncBar = 4;

// Written (once) by the first block whose grid barrier ran out of patience; lives in mapped pinned host memory so that the
// host can read it even while a kernel is still stuck.
struct TreeDebug {
    unsigned aborted;            // 0 = never; else 1 + number of aborts seen
    unsigned kernel;             // 0 k_tree, 1 k_tree_top, 2 k_front, 3 k_tree_df (a dependency wait, not a grid barrier)
    unsigned op;                 // 0 factor, 1 forward, 2 backward
    int phase;                   // height whose end-of-phase barrier gave up
    unsigned block, grid;
    unsigned seen, target;       // counter value last read, value waited for
    unsigned long long waited_ns;
    unsigned progress;           // (kernel << 24) | (op << 16) | height: written by block 0 at the start of every height
};

struct TreeView {
    FactorView F;
    const SnMeta* meta;          // [number of tasks], in task order
    const ChildRec* crec;        // [nsuper - #roots], grouped by parent
    const int* par_pos;          // indexed like sn_rows
    const double* image;         // factor: assembled panels to read instead of Lx (same layout), or null
    const double* dperm;         // ... and the per-iteration diagonal in pivot order to add to them ([batch][xw_stride]), or null
    double* Cx;                  // [batch][cx_stride] contribution blocks
    int64_t cx_stride;
    double* cvec;                // [batch][cv_stride] forward-sweep contribution vectors, indexed like sn_rows
    int64_t cv_stride;
    double* xw;                  // [batch][xw_stride] permuted right-hand side / solution
    int64_t xw_stride;
    unsigned* bar;               // grid barrier counter of THIS launch (slot 4 of its first height's sync record, zeroed
                                 // by the host together with the tickets: host and device cannot drift apart)
    int* tickets;                // [n_levels][kSyncStride] per height: 4 task tickets + 1 barrier counter (zeroed by the host before a launch)
    unsigned* abort_flag;        // device: set when a barrier gave up; every spinner polls it and leaves the kernel
    TreeDebug* dbg;              // mapped pinned host memory: what the first block that gave up saw
    long long spin_budget_ns;    // how long one block waits at one barrier before it gives up
    const int* phase_cls;        // [n_levels][5] first task of class 256 / 128 / 64 / 32, end of the phase
    int B;
    int slice[4], cap[4], maxh[4];   // per class: doubles of shared memory per group, of which panel, of which rows
    int maxw;
    int dbg_sn[8];               // MADIPM_TREE_TIMING builds: supernodes whose stage clocks are printed
};
__device__ __forceinline__ bool dbg_hit(const TreeView& T, int sn) {
    bool hit = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) hit |= T.dbg_sn[k] == sn;
    return hit;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// barrier of one task group: a warp, the whole block (256 threads in k_tree, 1024 in k_tree_top), or a named
// barrier for the 128- and 64-thread groups (ids 1-2 and 3-6; barrier 0 is __syncthreads)
template <int G>
__device__ __forceinline__ void gsync() {
    if (G == 32) __syncwarp();
    else if (G >= kTreeThreads) __syncthreads();
    else named_barrier((G == 128 ? 1 : 3) + static_cast<int>(threadIdx.x) / G, G);
}

// barrier among the first `nact` threads (a multiple of 32) of a task group: the warps of a group that still own rows of
// the panel in dense_ldl.  The whole group -> its own barrier; one warp -> __syncwarp; in between -> an auxiliary named
// barrier (7-8 for the two 128-thread groups of a block, 9 for a 256- or 1024-thread group; 64-thread groups never need one).
template <int G>
__device__ __forceinline__ void gsync_active(int nact) {
    if (G == 32 || nact <= 32) __syncwarp();
    else if (nact >= G) gsync<G>();
    else named_barrier(G == 128 ? 7 + static_cast<int>(threadIdx.x) / G : 9, nact);
}

// offset of column b inside a packed lower triangle of order nb, shifted so that [a] addresses row a >= b
__device__ __forceinline__ int packed_col(int b, int nb) { return b * nb - ((b * (b + 1)) >> 1); }

// packed lower triangle of order nb as one flat array: first entry of column b, and the column of entry e
// (float square root, exact after one correction step: the discriminant is an integer below 2^24)
__device__ __forceinline__ int tri_colstart(int b, int nb) { return b * nb - ((b * (b - 1)) >> 1); }
__device__ __forceinline__ int tri_col(int e, int nb) {
    const int two = 2 * nb + 1;
    int b = static_cast<int>((static_cast<float>(two) - sqrtf(static_cast<float>(two * two - 8 * e))) * 0.5f);
    b = max(0, min(b, nb - 1));
    if (tri_colstart(b + 1, nb) <= e) ++b;
    else if (tri_colstart(b, nb) > e) --b;
    return b;
}
constexpr int kFlat = 4;         // contribution-block entries a thread keeps in flight

template <class T>
__device__ __forceinline__ T ldg_rec(const T* p) {        // 32- or 48-byte record through the read-only path (16-byte loads)
    static_assert(sizeof(T) % 16 == 0 && sizeof(T) <= 48, "record size");
    union { int4 v[sizeof(T) / 16]; T t; } u;
#pragma unroll
    for (int k = 0; k < static_cast<int>(sizeof(T) / 16); ++k) u.v[k] = __ldg(reinterpret_cast<const int4*>(p) + k);
    return u.t;
}

// stage the child records [c0 + base, c0 + base + n) of a task into shared memory (n <= kChildChunk)
template <int G>
__device__ __forceinline__ int stage_children(const TreeView& T, ChildRec* cs, int c0, int c1, int base, int t) {
    const int n = min(kChildChunk, c1 - c0 - base);
    if (base > 0) gsync<G>();                              // everyone is done with the previous chunk
    if (t < n) cs[t] = ldg_rec(T.crec + c0 + base + t);
    gsync<G>();
    return n;
}

// ---- dense LDL' of an h x w panel in shared memory (ld = h), left-looking, blocked by 4 columns ------------------
// The panel is small and every task is latency-bound on its pivot chain, so the loop is organised to keep that
// chain short rather than to save flops.  Thread t owns row t (rows t, t + G, ... for tall panels).  Per block of
// four columns:
//   1. each row subtracts the contributions of all columns to the left from its four entries (registers; the
//      operand U = L D of the pivot rows is kept in the otherwise unused upper triangle of the diagonal block, so
//      the inner loop is one strided own-row load, four broadcast loads and four independent FMAs),
//   2. the four pivot rows publish their entries, one barrier,
//   3. every thread factorises the 4 x 4 diagonal block redundantly in registers and solves its own row against
//      it (no communication; reciprocals by rcp.approx + two Newton steps), writes L (pivot rows also U), one barrier.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

#ifndef LDL_PROBE
#define LDL_PROBE(k)
#endif
template <int G>
__device__ __forceinline__ void dense_ldl(double* Ps, int h, int w, int* fail, int t, double growth = 0.0, double tiny = 0.0) {
    constexpr int NB = 4;
    for (int jb = 0; jb < w; jb += NB) {
        LDL_PROBE(7);
        if (jb + (t & ~31) >= h) break;                   // this warp owns no row any more (nor will it later): it waits at the end
        const int nact = min(G, (h - jb + 31) & ~31);     // threads of the warps that still own rows
        const int bw = min(NB, w - jb);
        const int o1 = bw > 1 ? h : 0, o2 = bw > 2 ? 2 * h : 0, o3 = bw > 3 ? 3 * h : 0;   // missing columns alias column 0
        // 1. left-looking update of this block's entries of my rows (the first row stays in registers)
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const double* U = Ps + jb * h;                     // U[k + c * h] = L[jb + c][k] * d_k for k < jb (upper triangle)
#pragma unroll 1
        for (int i = jb + t; i < h; i += G) {
            const double* li = Ps + i;
            double* row = Ps + i + jb * h;
            double c0 = row[0], c1 = row[o1], c2 = row[o2], c3 = row[o3];
#pragma unroll 4
            for (int k = 0; k < jb; ++k) {
                const double l = li[k * h];
                c0 = fma(-l, U[k], c0);
                c1 = fma(-l, U[k + o1], c1);
                c2 = fma(-l, U[k + o2], c2);
                c3 = fma(-l, U[k + o3], c3);
            }
            if (i == jb + t) { a0 = c0; a1 = c1; a2 = c2; a3 = c3; }
            if (i < jb + bw || i != jb + t) {              // pivot rows publish the diagonal block; further rows of a
                row[0] = c0;                               // tall panel park their values (entries right of the diagonal
                if (bw > 1) row[h] = c1;                   // are scratch: the upper triangle is rewritten in step 3)
                if (bw > 2) row[2 * h] = c2;
                if (bw > 3) row[3 * h] = c3;
            }
        }
        LDL_PROBE(0);
        gsync_active<G>(nact);
        LDL_PROBE(1);
        // 2. diagonal block (identity-padded when fewer than 4 columns are left), redundantly in every thread
        const double* Db = Ps + jb * (h + 1);
        const double d0 = Db[0];
        double a10 = bw > 1 ? Db[1] : 0.0, d1 = bw > 1 ? Db[h + 1] : 1.0;
        double a20 = bw > 2 ? Db[2] : 0.0, a21 = bw > 2 ? Db[h + 2] : 0.0, d2 = bw > 2 ? Db[2 * h + 2] : 1.0;
        double a30 = bw > 3 ? Db[3] : 0.0, a31 = bw > 3 ? Db[h + 3] : 0.0, a32 = bw > 3 ? Db[2 * h + 3] : 0.0,
               d3 = bw > 3 ? Db[3 * h + 3] : 1.0;
        const double r0 = fast_rcp(d0);
        const double l10 = a10 * r0, l20 = a20 * r0, l30 = a30 * r0;
        d1 = fma(-a10, l10, d1); a21 = fma(-a20, l10, a21); a31 = fma(-a30, l10, a31);
        const double r1 = fast_rcp(d1);
        const double l21 = a21 * r1, l31 = a31 * r1;
        d2 = fma(-a21, l21, fma(-a20, l20, d2)); a32 = fma(-a31, l21, fma(-a30, l20, a32));
        const double r2 = fast_rcp(d2);
        const double l32 = a32 * r2;
        LDL_PROBE(2);
        LDL_PROBE(3);
        gsync_active<G>(nact);                            // everyone has read the block before the pivot rows are rewritten
        LDL_PROBE(4);
        // 3. my rows against the block, one branch-free path for pivot rows and rows below: forward substitution
        //    gives v = l D; column c gets v_c / d_c left of the row's diagonal and v_c (= d) on it.  Rows that
        //    become pivot rows later also leave v in the upper triangle (entry (jb + c, i)).
        //    (d3 = v3 of the fourth pivot row: its reciprocal is taken by the rows below from their own copy.)
        const double d3f = fma(-a32, l32, fma(-a31, l31, fma(-a30, l30, d3)));
        const double r3 = fast_rcp(d3f);
#pragma unroll 1
        for (int i = jb + t; i < h; i += G) {
            double* row = Ps + i + jb * h;
            double* up = Ps + jb + i * h;
            const int ri = i - jb;                        // 0..3 for a pivot row
            double v0, v1, v2, v3;
            if (i == jb + t) { v0 = a0; v1 = a1; v2 = a2; v3 = a3; }
            else { v0 = row[0]; v1 = row[o1]; v2 = row[o2]; v3 = row[o3]; }
            v1 = fma(-v0, l10, v1);
            v2 = fma(-v1, l21, fma(-v0, l20, v2));
            v3 = fma(-v2, l32, fma(-v1, l31, fma(-v0, l30, v3)));
            row[0] = ri == 0 ? v0 : v0 * r0;
            if (bw > 1 && ri >= 1) row[h] = ri == 1 ? v1 : v1 * r1;
            if (bw > 2 && ri >= 2) row[2 * h] = ri == 2 ? v2 : v2 * r2;
            if (bw > 3 && ri >= 3) row[3 * h] = ri == 3 ? v3 : v3 * r3;
            if (i < w) {
                if (ri > 0) up[0] = v0;
                if (bw > 1 && ri > 1) up[1] = v1;
                if (bw > 2 && ri > 2) up[2] = v2;
                if (bw > 3 && ri > 3) up[3] = v3;
            }
        }
        LDL_PROBE(5);
        gsync_active<G>(nact);
        LDL_PROBE(6);
    }
    gsync<G>();                                           // the warps that ran out of rows rejoin here
    // pivots: a zero / non-finite one poisons everything after it, so one pass over D at the end is enough
    bool bad = false;
    for (int k = t; k < w; k += G) bad |= bad_pivot(Ps[k * (h + 1)], tiny);
    // A pivot of roundoff size (dependent rows, delta_c = 0) is not zero, so the test above lets it through and the factor is
    // garbage; without pivoting the symptom is element growth: |L_ik| = |a_ik / d_k| beyond anything regularisation explains
    // (delta = 1e-8 against entries of order one gives 1e8).  One pass over the strictly lower part of the panel.
    if (growth > 0.0) {
        for (int k = 0; k < w; ++k)
            for (int i = k + 1 + t; i < h; i += G) bad |= !(fabs(Ps[i + k * h]) <= growth);
    }
    if (bad) atomicExch(fail, 1);
}

// ---- numeric factorisation of one small supernode ------------------------------------------------------
// shared: [cap doubles panel][maxh doubles (unused by the blocked LDL')][maxw doubles D][kChildChunk child records]
template <int G>
__device__ __forceinline__ void tree_forward_core(const TreeView& T, const SnMeta& m, int inst, const double* Ps, double* su, double* sy,
                                                  ChildRec* cs, int t);

// FWD: the task also does the forward-sweep step of its supernode for the right-hand side waiting in xw (the predictor's:
// it depends on the iterate only, /root/reference/src/kernels.jl:21-41), with the freshly factorised panel still in shared
// memory -- one of the four triangular sweeps of an iteration costs no panel traffic and no extra heights.
template <int G, bool FWD>
__device__ __forceinline__ void tree_factor_task(const TreeView& T, const SnMeta& m, int inst, double* sm, int cap, int maxh) {
    const FactorView& F = T.F;
    constexpr int NW = G / 32;
    const int t = threadIdx.x % G;
    const int lane = t & 31;
    const int gw = t >> 5;
    const int c0 = m.c0, c1 = m.c1;
    const int w = m.w, h = m.h, nb = h - w;
    double* Lx = F.Lx + inst * F.lx_stride;
    double* P = Lx + m.xoff;
    double* Cx = T.Cx + inst * T.cx_stride;
    double* Ps = sm;
    double* ucol = Ps + cap;
    double* dk = ucol + maxh;
    ChildRec* cs = reinterpret_cast<ChildRec*>(dk + ((T.maxw + 1) & ~1));
    const int np = h * w;
    const int nchunk = (np + 1) >> 1;
#ifdef MADIPM_TREE_TIMING
    long long tt[10];
#define TT(k) do { if (t == 0 && dbg_hit(T, m.sn)) tt[k] = clock64(); } while (0)
#else
#define TT(k)
#endif
    TT(0);
    gsync<G>();                                           // the previous task of this group is done with the scratch
    const double* Psrc = T.image ? T.image + inst * F.lx_stride + m.xoff : P;
    for (int c = t; c < nchunk; c += G) cp_async16(Ps + 2 * c, Psrc + 2 * c);
    // this iteration's diagonal of my columns (the image holds the constant part), added before the LDL'
    const double* dpm = T.dperm ? T.dperm + inst * T.xw_stride + m.fc : nullptr;
    const double dg0 = (dpm && t < w) ? __ldcg(dpm + t) : 0.0;
    const double dg1 = (dpm && t + G < w) ? __ldcg(dpm + t + G) : 0.0;
    // 1. extend-add: the parts of the children's contribution blocks that land in this panel
    bool waited = false;
    for (int base = 0; base < c1 - c0; base += kChildChunk) {
        const int n = stage_children<G>(T, cs, c0, c1, base, t);
        if (!waited) { cp_async_wait_all(); gsync<G>(); waited = true; }
        for (int ci = 0; ci < n; ++ci) {
            const ChildRec r = cs[ci];
            if (!r.pull) continue;                        // that child scattered its block itself
            // The block is one flat array (columns b < nbp, the ones that land in this panel, come first): thread t
            // takes entries t, t + G, ... with kFlat loads in flight, whatever the shape of the triangle.
            const int* pos = T.par_pos + r.below;
            const double* Cc = Cx + r.coff;
            const int e1 = tri_colstart(r.nbp, r.nb);
            for (int e0 = t; e0 < e1; e0 += kFlat * G) {
                int dsti[kFlat];
                double v[kFlat];
#pragma unroll
                for (int u = 0; u < kFlat; ++u) {
                    const int e = e0 + u * G;
                    dsti[u] = -1;
                    v[u] = 0.0;
                    if (e < e1) {
                        v[u] = __ldcg(Cc + e);
                        const int b = tri_col(e, r.nb);
                        dsti[u] = __ldg(pos + b + e - tri_colstart(b, r.nb)) + __ldg(pos + b) * h;
                    }
                }
#pragma unroll
                for (int u = 0; u < kFlat; ++u)
                    if (dsti[u] >= 0) Ps[dsti[u]] += v[u];
            }
            gsync<G>();
        }
    }
    if (!waited) { cp_async_wait_all(); gsync<G>(); }
    if (dpm) {                                            // entry (k, k) is next touched by thread k (block 0) or after a barrier
        if (t < w) Ps[t * (h + 1)] += dg0;
        if (t + G < w) Ps[(t + G) * (h + 1)] += dg1;
        for (int k = t + 2 * G; k < w; k += G) Ps[k * (h + 1)] += __ldcg(dpm + k);
    }
    TT(1);
    // 2. dense LDL' of the panel: D on the diagonal of the top w x w block, unit-lower L elsewhere
    // (both end with a barrier of the whole group; the blocked right-looking routine needs h <= 2 G, which the task classes
    //  guarantee except for fronts taller than 512 rows in the 256-thread class)
    if (h <= 2 * G) dense_ldl_blk<G>(Ps, h, w, ucol, F.fail + inst, t, F.growth, F.tiny);
    else dense_ldl<G>(Ps, h, w, F.fail + inst, t, F.growth, F.tiny);
    TT(2);
    for (int c = t; c < nchunk; c += G) reinterpret_cast<double2*>(P)[c] = reinterpret_cast<const double2*>(Ps)[c];
    if (nb == 0) {
        if (FWD) tree_forward_core<G>(T, m, inst, Ps, ucol, dk, cs, t);      // (dense_ldl ended with a barrier of the group)
        return;
    }
    for (int k = t; k < w; k += G) dk[k] = Ps[k * (h + 1)];
    gsync<G>();
    TT(3);
    // 3. contribution block C = -L21 D L21' (packed lower triangle) on the FP64 tensor cores: the lower-triangular
    //    8 x 8 blocks of C are dealt to the warps; a block is w / 4 DMMA.8x8x4 with both operands read straight from
    //    the k-major panel in shared memory (A = rows of L21, B = rows of L21 scaled by D on the fly)
    double* Cs = Cx + m.coff;
    {
        const int fr = lane >> 2, fk = lane & 3;
        const int nb8 = (nb + 7) >> 3;
        int bi = 0, bj = 0;                               // block q of the row-major enumeration of the lower triangle
        __syncwarp();                                     // mma.sync.aligned below: the warp must be converged
        for (int q = 0, mine = gw; mine < nb8 * (nb8 + 1) / 2; mine += NW) {
            while (q < mine) { if (++bj > bi) { ++bi; bj = 0; } ++q; }
            const int ra = w + bi * 8 + fr, rb = w + bj * 8 + fr;
            const bool ain = ra < h, bin = rb < h;
            double c0 = 0.0, c1 = 0.0;
            for (int k0 = 0; k0 < w; k0 += 4) {
                const int kk = k0 + fk;
                const bool kin = kk < w;
                const double a = (kin && ain) ? Ps[ra + kk * h] : 0.0;
                const double b = (kin && bin) ? Ps[rb + kk * h] * dk[kk] : 0.0;
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
            }
            const int arow = bi * 8 + fr, col0 = bj * 8 + 2 * fk;
            if (arow < nb) {
                if (arow >= col0) Cs[packed_col(col0, nb) + arow] = -c0;
                if (arow >= col0 + 1) Cs[packed_col(col0 + 1, nb) + arow] = -c1;
            }
        }
    }
    TT(4);
    // 4. the children's pass-through parts (rows and columns beyond this supernode's own columns)
    for (int base = 0; base < c1 - c0; base += kChildChunk) {
        int n = min(kChildChunk, c1 - c0 - base);
        if (c1 - c0 > kChildChunk) n = stage_children<G>(T, cs, c0, c1, base, t);   // else the records are still staged
        for (int ci = 0; ci < n; ++ci) {
            const ChildRec r = cs[ci];
            if (!r.pull || r.nbp == r.nb) continue;
            gsync<G>();                                   // C (or the previous child's additions) is in place
            const int* pos = T.par_pos + r.below;
            const double* Cc = Cx + r.coff;
            const int e1 = tri_colstart(r.nb, r.nb);
            for (int e0 = tri_colstart(r.nbp, r.nb) + t; e0 < e1; e0 += kFlat * G) {
                int off[kFlat];
                double v[kFlat];
#pragma unroll
                for (int u = 0; u < kFlat; ++u) {
                    const int e = e0 + u * G;
                    off[u] = -1;
                    v[u] = 0.0;
                    if (e < e1) {
                        v[u] = __ldcg(Cc + e);
                        const int b = tri_col(e, r.nb);
                        off[u] = packed_col(__ldg(pos + b) - w, nb) + __ldg(pos + b + e - tri_colstart(b, r.nb)) - w;
                    }
                }
#pragma unroll
                for (int u = 0; u < kFlat; ++u)
                    if (off[u] >= 0) v[u] += __ldcg(Cs + off[u]);
#pragma unroll
                for (int u = 0; u < kFlat; ++u)
                    if (off[u] >= 0) Cs[off[u]] = v[u];
            }
        }
    }
    TT(5);
#ifdef MADIPM_TREE_TIMING
    if (t == 0 && dbg_hit(T, m.sn))
        printf("factor sn %d G %d w %d h %d children %d | stage+extend %lld ldl %lld store %lld schur %lld pass %lld cycles\n", m.sn, G, w, h,
               c1 - c0, tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], tt[4] - tt[3], tt[5] - tt[4]);
#endif
    // 5. under a big parent: scatter C into the ancestors' panels (the big kernels read panels only)
    if (m.kind == 1) {
        gsync<G>();
        const int64_t r0 = m.rowptr;
        const int* rows = F.sn_rows + r0;
        for (int b = gw; b < nb; b += NW) {
            const int64_t cb = __ldg(F.row_colbase + r0 + w + b);
            const int64_t rb = __ldg(F.row_relbase + r0 + w + b);
            const double* col = Cs + packed_col(b, nb);
            for (int a = b + lane; a < nb; a += 32)
                atomicAdd(Lx + cb + target_pos(F, rb, w + a, rows, w + b), __ldcg(col + a));
        }
    }
    if (FWD) {
        gsync<G>();                                       // everyone is done with D (dk) and the staged child records
        tree_forward_core<G>(T, m, inst, Ps, ucol, dk, cs, t);
    }
}

// ---- forward sweep: y = L^-1 b, xw <- D^-1 y ---------------------------------------------------------------
// shared: [cap doubles panel][maxh doubles pass-through][maxw doubles y][kChildChunk child records]
template <int G>
__device__ __forceinline__ void tree_forward_task(const TreeView& T, const SnMeta& m, int inst, double* sm, int cap, int maxh) {
    const FactorView& F = T.F;
    const int t = threadIdx.x % G;
    const double* P = F.Lx + inst * F.lx_stride + m.xoff;
    double* Ps = sm;
    double* su = Ps + cap;
    double* sy = su + maxh;
    ChildRec* cs = reinterpret_cast<ChildRec*>(sy + ((T.maxw + 1) & ~1));
    const int nchunk = (m.h * m.w + 1) >> 1;
    gsync<G>();
    for (int c = t; c < nchunk; c += G) cp_async16(Ps + 2 * c, P + 2 * c);
    tree_forward_core<G>(T, m, inst, Ps, su, sy, cs, t);
}

// The forward step of one supernode on a panel that is (or is on its way to being, cp.async) in shared memory at Ps.
// Shared by the forward sweep proper and by the factor task with FWD: same code, same arithmetic, bit for bit.
template <int G>
__device__ __forceinline__ void tree_forward_core(const TreeView& T, const SnMeta& m, int inst, const double* Ps, double* su, double* sy,
                                                  ChildRec* cs, int t) {
    const FactorView& F = T.F;
    const int lane = t & 31;
    const int gw = t >> 5;
    const int c0 = m.c0, c1 = m.c1;
    const int w = m.w, h = m.h, nb = h - w, fc = m.fc;
    double* xw = T.xw + inst * T.xw_stride;
    double* cv = T.cvec + inst * T.cv_stride;
#ifdef MADIPM_TREE_TIMING
    long long tt[6];
#endif
    TT(0);
    for (int j = t; j < w; j += G) sy[j] = __ldcg(xw + fc + j);
    for (int i = t; i < nb; i += G) su[i] = 0.0;
    for (int base = 0; base < c1 - c0; base += kChildChunk) {
        const int n = stage_children<G>(T, cs, c0, c1, base, t);   // its barrier also publishes sy / su
        // The contribution vectors of all staged children are fetched in ONE pass (entry e of the concatenated
        // vectors goes to thread e), so their memory round trips overlap; they are then applied child by child,
        // which keeps the summation order fixed.
        int total = 0;
        for (int ci = 0; ci < n; ++ci) total += cs[ci].pull ? cs[ci].nb : 0;
        for (int e0 = 0; e0 < total; e0 += G) {
            int e = e0 + t, mine = -1, pa = 0;
            double v = 0.0;
            if (e < total) {
                for (int ci = 0; ci < n; ++ci) {
                    const int len = cs[ci].pull ? cs[ci].nb : 0;
                    if (e < len) { mine = ci; break; }
                    e -= len;
                }
                pa = __ldg(T.par_pos + cs[mine].below + e);
                v = __ldcg(cv + cs[mine].below + e);
            }
            for (int ci = 0; ci < n; ++ci) {
                if (mine == ci) {
                    if (pa < w) sy[pa] -= v;
                    else su[pa - w] += v;
                }
                gsync<G>();
            }
        }
    }
    TT(1);
    cp_async_wait_all();
    gsync<G>();
    TT(2);
    if (gw == 0) {                                        // unit-lower triangular solve, one warp
        __syncwarp();                                     // (full-mask shuffles below)
        if (w <= 64) {
            // y in registers (lane l holds y[l] and y[l + 32]); per column one shuffle and one FMA on the dependent
            // chain, the entries of L come from shared memory ahead of it
            double y0 = lane < w ? sy[lane] : 0.0, y1 = lane + 32 < w ? sy[lane + 32] : 0.0;
            const double* Pl = Ps + lane;
            const int w0 = min(w, 32);
#pragma unroll 4
            for (int j = 0; j < w0; ++j) {
                const double l0 = (lane > j && lane < w) ? Pl[j * h] : 0.0;
                const double l1 = lane + 32 < w ? Pl[32 + j * h] : 0.0;
                const double yj = __shfl_sync(0xffffffffu, y0, j);
                y0 = fma(-l0, yj, y0);
                y1 = fma(-l1, yj, y1);
            }
#pragma unroll 4
            for (int j = 32; j < w; ++j) {
                const double l1 = (lane + 32 > j && lane + 32 < w) ? Pl[32 + j * h] : 0.0;
                const double yj = __shfl_sync(0xffffffffu, y1, j - 32);
                y1 = fma(-l1, yj, y1);
            }
            if (lane < w) sy[lane] = y0;
            if (lane + 32 < w) sy[lane + 32] = y1;
        } else {
            for (int j = 0; j + 1 < w; ++j) {
                const double yj = sy[j];
                for (int i = j + 1 + lane; i < w; i += 32) sy[i] -= Ps[i + j * h] * yj;
                __syncwarp();
            }
        }
    }
    gsync<G>();
    TT(3);
    const bool pull = m.kind == 0;
    const int* rows = F.sn_rows + m.rowptr + w;
    for (int i = t; i < nb; i += G) {
        double acc = su[i];
        const double* Li = Ps + w + i;
#pragma unroll 4
        for (int j = 0; j < w; ++j) acc = fma(Li[j * h], sy[j], acc);
        if (pull) cv[m.rowptr + w + i] = acc;
        else atomicAdd(xw + __ldg(rows + i), -acc);
    }
    for (int j = t; j < w; j += G) xw[fc + j] = sy[j] / Ps[j * (h + 1)];
    TT(4);
#ifdef MADIPM_TREE_TIMING
    if (t == 0 && dbg_hit(T, m.sn))
        printf("forward sn %d G %d w %d h %d children %d | children %lld panel-wait %lld tri %lld matvec+store %lld cycles\n", m.sn, G, w, h,
               c1 - c0, tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], tt[4] - tt[3]);
#endif
}

// ---- backward sweep: x = L^-T xw -----------------------------------------------------------------------------
// shared: [cap doubles panel][maxh doubles x at the below rows][maxw doubles block solution]
template <int G>
__device__ __forceinline__ void tree_backward_task(const TreeView& T, const SnMeta& m, int inst, double* sm, int cap, int maxh) {
    const FactorView& F = T.F;
    constexpr int NW = G / 32;
    const int t = threadIdx.x % G;
    const int lane = t & 31;
    const int gw = t >> 5;
    const int w = m.w, h = m.h, nb = h - w, fc = m.fc;
    const double* P = F.Lx + inst * F.lx_stride + m.xoff;
    double* xw = T.xw + inst * T.xw_stride;
    double* Ps = sm;
    double* sx = Ps + cap;
    double* st = sx + maxh;
    const int nchunk = (h * w + 1) >> 1;
#ifdef MADIPM_TREE_TIMING
    long long tt[6];
#endif
    TT(0);
    gsync<G>();
    for (int c = t; c < nchunk; c += G) cp_async16(Ps + 2 * c, P + 2 * c);
    const int* rows = F.sn_rows + m.rowptr + w;
    for (int j = t; j < w; j += G) st[j] = __ldcg(xw + fc + j);
    for (int i = t; i < nb; i += G) sx[i] = __ldcg(xw + __ldg(rows + i));
    cp_async_wait_all();
    gsync<G>();
    TT(1);
    for (int j = gw; j < w; j += NW) {
        const double* Lj = Ps + w + j * h;
        double acc = 0.0;
        for (int i = lane; i < nb; i += 32) acc = fma(Lj[i], sx[i], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) st[j] -= acc;
    }
    gsync<G>();
    TT(2);
    if (gw == 0) {                                        // unit-upper triangular solve with L11', one warp
        __syncwarp();                                     // (full-mask shuffles below)
        if (w <= 64) {
            // x in registers (lane l holds x[l] and x[l + 32]); L11'[k][j] = L[j][k] = Ps[j + k * h]
            double x0 = lane < w ? st[lane] : 0.0, x1 = lane + 32 < w ? st[lane + 32] : 0.0;
            const double* Pk0 = Ps + lane * h;
            const double* Pk1 = Ps + (lane + 32) * h;
#pragma unroll 4
            for (int j = w - 1; j >= 32; --j) {
                const double l0 = Pk0[j];                                      // lane < 32 <= j < w
                const double l1 = lane + 32 < j ? Pk1[j] : 0.0;
                const double xj = __shfl_sync(0xffffffffu, x1, j - 32);
                x0 = fma(-l0, xj, x0);
                x1 = fma(-l1, xj, x1);
            }
#pragma unroll 4
            for (int j = min(w, 32) - 1; j > 0; --j) {
                const double l0 = lane < j ? Pk0[j] : 0.0;
                const double xj = __shfl_sync(0xffffffffu, x0, j);
                x0 = fma(-l0, xj, x0);
            }
            if (lane < w) st[lane] = x0;
            if (lane + 32 < w) st[lane + 32] = x1;
        } else {
            for (int j = w - 1; j > 0; --j) {
                const double xj = st[j];
                for (int k = lane; k < j; k += 32) st[k] -= Ps[j + k * h] * xj;
                __syncwarp();
            }
        }
    }
    gsync<G>();
    TT(3);
    for (int j = t; j < w; j += G) xw[fc + j] = st[j];
#ifdef MADIPM_TREE_TIMING
    if (t == 0 && dbg_hit(T, m.sn))
        printf("backward sn %d G %d w %d h %d | load %lld matvec %lld tri %lld cycles\n", m.sn, G, w, h, tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2]);
#endif
}

// ---- grid barrier (cooperative launch: every block is resident) -------------------------------------------
// A counter in global memory that the host zeroes before every launch; barrier k of a launch waits for k * gridDim.x
// arrivals.  The spin is BOUNDED: a block that waits longer than T.spin_budget_ns raises T.abort_flag and records what it
// saw in T.dbg (host-visible); every spinner polls the flag, and an aborted barrier makes the whole grid leave the kernel.
// The host turns the flag into MADIPM_INTERNAL at its next synchronisation -- a desynchronised grid is an error message,
// not a GPU pinned at 100 % forever.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool grid_barrier(const TreeView& T, unsigned& target, int kernel, int op, int ph) {
    __shared__ int s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(T.bar, 1u);
        unsigned v, polls = 0;
        unsigned long long t0 = 0;
        int ok = 1;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(T.bar) : "memory");
            if (static_cast<int>(v - target) >= 0) break;
            if ((++polls & 1023u) == 0) {
                unsigned ab;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ab) : "l"(T.abort_flag) : "memory");
                if (ab) { ok = 0; break; }
                const unsigned long long now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > static_cast<unsigned long long>(T.spin_budget_ns)) {
                    if (atomicExch(T.abort_flag, 1u) == 0u) {
                        TreeDebug* g = T.dbg;
                        g->kernel = kernel; g->op = op; g->phase = ph; g->block = blockIdx.x; g->grid = gridDim.x;
                        g->seen = v; g->target = target; g->waited_ns = now - t0;
                        __threadfence_system();
                        g->aborted = g->aborted + 1u;
                        __threadfence_system();
                    }
                    ok = 0;
                    break;
                }
            }
        }
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}

template <int OP, int G>
__device__ __forceinline__ void tree_task(const TreeView& T, const SnMeta& m, int inst, double* sm, int cap, int maxh) {
    if (OP == 0) tree_factor_task<G, false>(T, m, inst, sm, cap, maxh);
    else if (OP == 3) tree_factor_task<G, true>(T, m, inst, sm, cap, maxh);
    else if (OP == 1) tree_forward_task<G>(T, m, inst, sm, cap, maxh);
    else tree_backward_task<G>(T, m, inst, sm, cap, maxh);
}

// ---- dataflow scheduling (k_tree_df): no barrier between two heights --------------------------------------------------
// The ncu stall samples of k_tree say where its time goes: blocks that ran out of tasks of a height wait at the grid barrier
// for the slowest front of that height.  In the dataflow variant a task waits for ITS OWN
// dependencies instead: one counter per (instance, task) in global memory, zeroed by the host before the launch.
//   * factor / forward (children -> parent): a finished task adds 1 to its parent's counter (all stores of the group, group
//     barrier, __threadfence, atomicAdd by one thread -- the pattern of the grid barrier); a task starts when its counter has
//     reached its number of children (ld.acquire.gpu by one thread, group barrier, then the usual L2 reads).
//   * backward (parent -> children): a finished task sets its own counter; a task starts when its parent's is set.  Every
//     ancestor's part of x is visible then, because the parent waited for ITS parent before it ran (causality is transitive).
// Tickets are still handed out height by height, class by class, and a block moves to the next height only after it has seen
// the ticket counters of all classes of this height exhausted: every task a ticket holder waits for is held by a resident
// group (cooperative launch) that waits only for tasks of lower heights -- no cycle, no deadlock.  Nothing about the
// arithmetic changes (same tasks, same summation order): the results are bit-identical to the barrier schedule (tested).
// The spin is bounded like the grid barrier's: on a timeout or a raised abort flag the group skips the task WITHOUT
// signalling, everybody downstream gives up at once (they poll the flag), the tickets drain, the kernel ends, and the host
// reports MADIPM_INTERNAL at its next synchronisation.
struct DfView {
    int* dep;                    // [batch][ntasks] dependency counters of this launch (zeroed by the host)
    const int* task_parent;      // [ntasks] task index of the parent supernode, -1 for a root
    int ntasks;
    int df_end;                  // tasks [0, df_end) belong to this launch; parents beyond it are left to the next launch
};

__device__ __forceinline__ bool df_spin(const TreeView& T, const int* addr, int need, int op, int ph) {
    unsigned polls = 0;
    unsigned long long t0 = 0;
    for (;;) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
        if (v >= need) return true;
        if ((polls++ & 1023u) == 0) {
            unsigned ab;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(ab) : "l"(T.abort_flag) : "memory");
            if (ab) return false;
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > static_cast<unsigned long long>(T.spin_budget_ns)) {
                if (atomicExch(T.abort_flag, 1u) == 0u) {
                    TreeDebug* g = T.dbg;
                    g->kernel = 3; g->op = op; g->phase = ph; g->block = blockIdx.x; g->grid = gridDim.x;
                    g->seen = static_cast<unsigned>(v); g->target = static_cast<unsigned>(need); g->waited_ns = now - t0;
                    __threadfence_system();
                    g->aborted = g->aborted + 1u;
                    __threadfence_system();
                }
                return false;
            }
        }
    }
}

// one group of G threads drains the tasks of its class in phase ph (ticket counter), then returns
template <int OP, int G, int CLS, bool DF>
__device__ __forceinline__ void run_class(const TreeView& T, const DfView& D, double* smem, int ph) {
    const int first = __ldg(T.phase_cls + ph * 5 + CLS);
    const int n = __ldg(T.phase_cls + ph * 5 + CLS + 1) - first;
    if (n == 0) return;
    const int64_t total = static_cast<int64_t>(n) * T.B;
    double* sm = smem + (static_cast<int>(threadIdx.x) / G) * T.slice[CLS];
    volatile int* slot = reinterpret_cast<volatile int*>(sm + T.slice[CLS] - 2);
    const int t = threadIdx.x % G;
    for (;;) {
        gsync<G>();                                       // the previous ticket has been read by everyone
        if (t == 0) *slot = atomicAdd(T.tickets + ph * kSyncStride + CLS, 1);
        gsync<G>();
        const int64_t q = *slot;
        if (q >= total) break;
        const int inst = static_cast<int>(q / n);
        if (DF) {
            const int tq = first + static_cast<int>(q - static_cast<int64_t>(inst) * n);
            const SnMeta m = ldg_rec(T.meta + tq);
            int* dep = D.dep + static_cast<int64_t>(inst) * D.ntasks;
            const int par = __ldg(D.task_parent + tq);
            const bool par_here = par >= 0 && par < D.df_end;
            const int need = OP == 2 ? (par_here ? 1 : 0) : m.c1 - m.c0;
            if (need > 0) {                               // (uniform over the group: m and par are)
                if (t == 0) slot[1] = df_spin(T, dep + (OP == 2 ? par : tq), need, OP, ph) ? 1 : 0;
                gsync<G>();
                if (slot[1] == 0) continue;               // aborted: skip the task, do not signal; the loop's first barrier
            }                                             // keeps slot[1] from being rewritten before everyone has read it
            tree_task<OP, G>(T, m, inst, sm, T.cap[CLS], T.maxh[CLS]);
            gsync<G>();                                   // every store of the task has been issued
            if (t == 0) {
                __threadfence();
                if (OP == 2) { if (m.c1 > m.c0) atomicExch(dep + tq, 1); }      // (nobody waits for a leaf)
                else if (par_here) atomicAdd(dep + par, 1);
            }
        } else {
            tree_task<OP, G>(T, ldg_rec(T.meta + first + static_cast<int>(q - static_cast<int64_t>(inst) * n)), inst, sm, T.cap[CLS], T.maxh[CLS]);
        }
    }
}

// OP: 0 factor, 1 forward, 2 backward, 3 factor + forward step of the right-hand side in xw.  Heights ph_begin, ph_begin + step, ... (ph_end excluded).
#ifndef MADIPM_TREE_BLOCKS
#define MADIPM_TREE_BLOCKS 4
#endif
template <int OP>
__global__ void __launch_bounds__(kTreeThreads, MADIPM_TREE_BLOCKS)
k_tree(TreeView T, int ph_begin, int ph_end, int step) {
    extern __shared__ double2 tree_smem2[];
    double* smem = reinterpret_cast<double*>(tree_smem2);
    unsigned target = 0;
    const DfView D{};
    for (int ph = ph_begin; ph != ph_end; ph += step) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(&T.dbg->progress) = (0u << 24) | (OP << 16) | ph;
        run_class<OP, 256, 0, false>(T, D, smem, ph);
        run_class<OP, 128, 1, false>(T, D, smem, ph);
        run_class<OP, 64, 2, false>(T, D, smem, ph);
        run_class<OP, 32, 3, false>(T, D, smem, ph);
        if (ph + step != ph_end && !grid_barrier(T, target, 0, OP, ph)) return;
    }
}

// the same walk without the barrier between two heights: every task waits for its own dependencies (see DfView)
template <int OP>
__global__ void __launch_bounds__(kTreeThreads, MADIPM_TREE_BLOCKS)
k_tree_df(TreeView T, DfView D, int ph_begin, int ph_end, int step) {
    extern __shared__ double2 tree_smem2[];
    double* smem = reinterpret_cast<double*>(tree_smem2);
    for (int ph = ph_begin; ph != ph_end; ph += step) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(&T.dbg->progress) = (3u << 24) | (OP << 16) | ph;
        run_class<OP, 256, 0, true>(T, D, smem, ph);
        run_class<OP, 128, 1, true>(T, D, smem, ph);
        run_class<OP, 64, 2, true>(T, D, smem, ph);
        run_class<OP, 32, 3, true>(T, D, smem, ph);
    }
}

// the heights near the root: at most one task per SM, one 1024-thread block per task
constexpr int kTopThreads = 1024;
template <int OP>
__global__ void __launch_bounds__(kTopThreads, 1)
k_tree_top(TreeView T, int ph_begin, int ph_end, int step) {
    extern __shared__ double2 tree_smem2[];
    double* smem = reinterpret_cast<double*>(tree_smem2);
    unsigned target = 0;
    for (int ph = ph_begin; ph != ph_end; ph += step) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(&T.dbg->progress) = (1u << 24) | (OP << 16) | ph;
        const int first = __ldg(T.phase_cls + ph * 5);
        const int n = __ldg(T.phase_cls + ph * 5 + 4) - first;
        const int64_t total = static_cast<int64_t>(n) * T.B;
        for (int64_t q = blockIdx.x; q < total; q += gridDim.x) {
            const int inst = static_cast<int>(q / n);
            tree_task<OP, kTopThreads>(T, ldg_rec(T.meta + first + static_cast<int>(q - static_cast<int64_t>(inst) * n)), inst, smem,
                                       T.cap[0], T.maxh[0]);
        }
        if (ph + step != ph_end && !grid_barrier(T, target, 1, OP, ph)) return;
    }
}

}  // namespace madipm
