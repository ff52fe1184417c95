// Dense LDL' of one h x w front held in shared memory, rows in REGISTERS, right-looking, one barrier per column.
//
// Why: a front near the root of the elimination tree is one task on one SM and nothing else can run until it is done
// (the height is a phase), so what counts is the length of the dependent chain per column, not flops.  Thread t owns row t
// of the panel and keeps the part of it that is still to be eliminated in registers.  Column j:
//   * everybody reads the published pivot column (d_j and u_r = a_{j+r,j}, r >= 1: the entries of the pivot block below the
//     diagonal, before scaling) from a small shared buffer, indexed RELATIVE to the pivot (broadcast 16-byte loads at
//     compile-time offsets),
//   * l_ij = a_ij / d_j goes to the panel, then FIRST the entry of the next column, a_i,j+1 -= l_ij u_1, which is
//     published at once for step j + 1 (look-ahead), then the rest of the row, a_i,j+k -= l_ij u_k,
//   * the register row SHIFTS by one (the FMA for a_k writes register k - 1), so the loop body has compile-time register
//     indices while j is a run-time loop variable: the code is one compact loop per register width (W, W - 8, ..., 8: the
//     width follows the shrinking trailing row), a few hundred instructions in all.  (A fully unrolled j loop is 50 KB of
//     straight-line code per width that is executed once per front: measured 14 k cycles with a warm instruction cache and
//     30 k inside the factorisation kernel, where it never is.)
//   * one barrier (the buffer is double-buffered by column parity).
// The blocked left-looking routine (dense_ldl in kernels_tree.cuh) keeps the row in shared memory and needs ~600 cycles per
// column on the same fronts (~1100 inside k_tree_top with its 32 warps per barrier).
#pragma once
#include <cuda_runtime.h>

namespace madipm {

__device__ __forceinline__ double rr_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

// barrier among the first `n` threads (multiple of 32) of the group that uses named barrier `id`
__device__ __forceinline__ void rr_bar(int id, int n) {
    if (n <= 32) __syncwarp();
    else {
#ifdef MADIPM_ALIGNED_BARRIERS
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
#else
        __syncwarp();
        asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(n) : "memory");     // non-aligned on purpose: see named_barrier in kernels_factor.cuh
        __syncwarp();
#endif
    }
}

constexpr int kRrLd = 64 + 2;           // leading dimension of the two column buffers (doubles)

struct RrState {
    double* Ps;
    double* buf;
    int h, w, t, nact, bar_id;
    double growth, tiny;
    bool bad;
};

// columns J, J + 1, ... with a register row of width W (a[k] = entry (t, J + k)), while more than W - 8 columns remain
template <int WMAX, int W>
struct rr_phase {
    static __device__ __forceinline__ void run(double (&a)[WMAX], RrState& S, int& J) {
        const int w = S.w, h = S.h, t = S.t;
#pragma unroll 1
        while (J < w && (W == 8 || w - J > W - 8)) {
            const double* u = S.buf + (J & 1) * kRrLd;
            double* un = S.buf + ((J + 1) & 1) * kRrLd;
            const double2 u01 = *reinterpret_cast<const double2*>(u);
            const double d = u01.x;
            S.bad |= !(fabs(d) > S.tiny && fabs(d) < 1.0e300);
            const double l = a[0] * rr_rcp(d);
            const double a1 = fma(-l, u01.y, a[1]);
            if (t == J) S.Ps[J * (h + 1)] = d;                       // D stays on the diagonal
            if (t > J) {
                if (t < w) un[t - J - 1] = a1;                       // rows J+1 .. w-1 publish column J+1 (row J+1: its pivot)
                if (t < h) {
                    S.Ps[t + J * h] = l;                             // unit-lower factor
                    S.bad |= S.growth > 0.0 && !(fabs(l) <= S.growth);
                }
            }
            a[0] = a1;
#pragma unroll
            for (int k = 2; k + 1 < W; k += 2) {
                const double2 uu = *reinterpret_cast<const double2*>(u + k);
                a[k - 1] = fma(-l, uu.x, a[k]);
                a[k] = fma(-l, uu.y, a[k + 1]);
            }
            a[W - 1] = 0.0;
            ++J;
            rr_bar(S.bar_id, S.nact);
        }
        if constexpr (W > 8) rr_phase<WMAX, W - 8>::run(a, S, J);
    }
};

// Ps: panel, column-major, leading dimension h; on return D on the diagonal, unit-lower L below (the strict upper triangle
// of the pivot block is left alone).  buf: 2 * kRrLd doubles of shared memory, 16-byte aligned.  t: index of the thread in
// its group; the group has at least ((h + 31) & ~31) threads; threads with t >= that return at once (false).  bar_id: a named
// barrier reserved for this group.  Requires w <= WMAX <= 64 (WMAX a multiple of 8), h >= w.  Returns true where this
// thread saw a zero / tiny (|d| <= tiny) / non-finite pivot or (growth > 0) a factor entry of its row above `growth` in magnitude.
template <int WMAX>
__device__ __forceinline__ bool dense_ldl_rr(double* __restrict__ Ps, int h, int w, double* __restrict__ buf, int t, int bar_id,
                                             double growth = 0.0, double tiny = 0.0) {
    static_assert(WMAX % 8 == 0 && WMAX <= 64, "WMAX: multiple of 8, at most 64");
    const int nact = (h + 31) & ~31;
    if (t >= nact) return false;
    const bool live = t < h;
    double a[WMAX];
#pragma unroll
    for (int k = 0; k < WMAX; ++k) a[k] = (live && k < w && k <= t) ? Ps[t + k * h] : 0.0;
    // column 0, relative to its pivot; the rest of both buffers is zero (entries past the pivot block multiply padding only)
    for (int k = t; k < 2 * kRrLd; k += nact) buf[k] = 0.0;
    rr_bar(bar_id, nact);
    if (live && t < w) buf[t] = a[0];
    rr_bar(bar_id, nact);
    RrState S{Ps, buf, h, w, t, nact, bar_id, growth, tiny, false};
    int J = 0;
    rr_phase<WMAX, WMAX>::run(a, S, J);
    return S.bad;
}

}  // namespace madipm
