// Dense LDL' of one h x w front held in shared memory, rows in REGISTERS, right-looking, one barrier per column.
//
// Why: a front near the root of the elimination tree is one task on one SM and nothing else can run until it is done
// (the height is a phase), so what counts is the length of the dependent chain per column, not flops.  Thread t owns row t
// of the panel and keeps its w entries in registers (compile-time bound WMAX, fully unrolled).  Column j:
//   * everybody reads the published column  u_k = a_kj (k > j, the entries of the pivot block below the diagonal, before
//     scaling) and d_j from a small shared buffer (broadcast loads),
//   * l_ij = a_ij / d_j, then FIRST the entry of the next column, a_i,j+1 -= l_ij u_j+1, which is published at once for
//     step j + 1 (look-ahead), then the rest of the row, a_ik -= l_ij u_k,
//   * one barrier (the buffer is double-buffered by column parity).
// The chain per column is LDS -> rcp -> mul -> fma -> STS -> barrier, about 150 cycles; the blocked left-looking routine
// (dense_ldl in kernels_tree.cuh) needs about 900 per column on the same fronts because its inner loops run over
// shared-memory operands.
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
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// Column steps j = J .. WMAX-1 as a compile-time recursion: every index into the register row is a constant (a runtime
// loop, or an unroll the compiler declines, would send the row to local memory).  Columns w .. WMAX-1 are zero padding
// (a_k = 0 and u_k = 0 there), so the row update runs unpredicated over the compile-time width.
template <int WMAX, int J>
struct rr_steps {
    static __device__ __forceinline__ void run(double (&a)[WMAX], double* __restrict__ buf, int w, int t, int nact, int bar_id) {
        constexpr int LD = WMAX + 2;
        if (J >= w) return;
        const double* u = buf + (J & 1) * LD;
        double* un = buf + ((J + 1) & 1) * LD;
        const double l = a[J] * rr_rcp(u[J]);
        constexpr int J1 = (J + 1 < WMAX) ? J + 1 : J;
        if (J + 1 < WMAX) {
            a[J1] = fma(-l, u[J1], a[J1]);
            if (t > J && t < w) un[t] = a[J1];                            // rows J+1 .. w-1 publish column J+1
        }
        // the rest of the row: u in aligned pairs (one 16-byte broadcast load per two columns)
        constexpr int K0 = J + 2;
        constexpr int KA = (K0 + 1) & ~1;                                 // first even index >= K0
        if (K0 < KA && K0 < WMAX) a[K0 < WMAX ? K0 : 0] = fma(-l, u[K0 < WMAX ? K0 : 0], a[K0 < WMAX ? K0 : 0]);
#pragma unroll
        for (int k = KA; k + 1 < WMAX; k += 2) {
            const double2 uu = *reinterpret_cast<const double2*>(u + k);
            a[k] = fma(-l, uu.x, a[k]);
            a[k + 1] = fma(-l, uu.y, a[k + 1]);
        }
        if (t > J) a[J] = l;                                              // unit-lower factor; row J keeps d_J
        rr_bar(bar_id, nact);
        rr_steps<WMAX, J + 1>::run(a, buf, w, t, nact, bar_id);
    }
};
template <int WMAX>
struct rr_steps<WMAX, WMAX> {
    static __device__ __forceinline__ void run(double (&)[WMAX], double*, int, int, int, int) {}
};

// Ps: panel, column-major, leading dimension h; on return D on the diagonal, unit-lower L below (strict upper triangle of
// the pivot block: untouched).  buf: 2 * (WMAX + 2) doubles of shared memory, 16-byte aligned.  t: index of the thread in
// its group; the group has at least ((h + 31) & ~31) threads; threads with t >= that return at once.  bar_id: a named
// barrier reserved for this group.  Requires w <= WMAX (even), h >= w.  Pivots are not checked here (the caller looks at D).
template <int WMAX>
__device__ __forceinline__ void dense_ldl_rr(double* __restrict__ Ps, int h, int w, double* __restrict__ buf, int t, int bar_id) {
    static_assert(WMAX % 2 == 0, "WMAX must be even");
    constexpr int LD = WMAX + 2;
    const int nact = (h + 31) & ~31;
    if (t >= nact) return;
    const bool live = t < h;
    double a[WMAX];
#pragma unroll
    for (int k = 0; k < WMAX; ++k) a[k] = (live && k < w && k <= t) ? Ps[t + k * h] : 0.0;
    // zero padding of both column buffers beyond w, then column 0
    for (int k = w + t; k < 2 * LD; k += nact)
        if (k < LD || k >= LD + w) buf[k] = 0.0;
    if (live && t < w) buf[t] = a[0];
    rr_bar(bar_id, nact);
    rr_steps<WMAX, 0>::run(a, buf, w, t, nact, bar_id);
    if (live) {
#pragma unroll
        for (int k = 0; k < WMAX; ++k)
            if (k < w && k <= t) Ps[t + k * h] = a[k];
    }
}

}  // namespace madipm
