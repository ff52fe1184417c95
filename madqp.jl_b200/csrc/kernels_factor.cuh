// Shared pieces of the supernodal LDL' (north-star part (b)).
//
// Storage: supernode s owns an h x w column-major panel (ld = h) at Lx + xoff[s]; its first w
// rows are its own columns.  After factorisation the panel holds D on the diagonal of the top
// w x w block, unit-lower L below it.
//
// Small supernodes are factorised by the height-scheduled kernels of kernels_tree.cuh, big ones
// (panel too large for one block) by the multi-block kernels of kernels_big.cuh.  This header
// holds what both use: the view of the factor structure, the scatter tables that say where a
// Schur-complement entry lands in an ancestor's panel (tabulated at symbolic time, so a scatter
// has no searches), the assembly of the KKT values into the panels, and the inertia count.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace madipm {

struct FactorView {
    const int* sn_first;       // [nsuper+1]
    const int64_t* sn_rowptr;  // [nsuper+1]
    const int* sn_rows;
    const int64_t* sn_xoff;    // [nsuper+1]
    const int* sn_of;          // [N]
    const int64_t* row_colbase;   // indexed like sn_rows
    const int64_t* row_relbase;
    const int* rel;
    double* Lx;                // [batch][lx_stride]
    int64_t lx_stride;
    int* fail;                 // [batch]
    double growth;             // element-growth limit of the unpivoted LDL': |L_ik| above it is a breakdown (0 = no check)
    double tiny;               // |pivot| at or below it is a breakdown (0 = only exact zeros, the reference's semantics)
};

// Named barrier among `count` threads (a multiple of 32) of a thread block.
//
// This MUST be the non-aligned form.  PTX `bar.sync` is `barrier.sync.aligned`: every thread of a warp has to execute it
// together, and the result is undefined when the warp's lanes have not reconverged after divergent code (a data-dependent
// loop, an `if (t < n)` store) -- the compiler usually reconverges them in time, which is what made this a rare,
// timing- and register-allocation-dependent hang instead of a plain bug: a warp whose halves arrive separately is counted
// twice, the barrier opens early once and is then one arrival short for ever; the group never finishes its task, its block
// never reaches the grid barrier, the persistent kernel never returns.  That was the 30-minute hang of round 1's driver
// benchmark; it became deterministic with -DMADIPM_TREE_BLOCKS=3 (80 registers, different instruction schedule) and went away
// with `barrier.sync`, which counts threads individually.  -DMADIPM_ALIGNED_BARRIERS restores the old form (to reproduce).
// The __syncwarp() on both sides makes the warp arrive and leave CONVERGED: code after a barrier uses warp collectives
// (mma.sync.aligned, __shfl_sync with the full mask) whose convergence requirement is the same as bar.sync's.
__device__ __forceinline__ void named_barrier(int id, int count) {
#ifdef MADIPM_ALIGNED_BARRIERS
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#else
    __syncwarp();
    asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
    __syncwarp();
#endif
}

constexpr int64_t kRelShift = int64_t(1) << 40;
constexpr int64_t kRelSearch = INT64_MIN;

__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }

__device__ __forceinline__ int lower_bound_rows(const int* rows, int n, int key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (rows[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ bool bad_pivot(double d, double tiny = 0.0) { return !(fabs(d) > tiny && fabs(d) < 1.0e300); }

// position of rows[i] inside the target's row list
__device__ __forceinline__ int target_pos(const FactorView& F, int64_t rb, int i, const int* rows, int jj) {
    if (rb >= 0) return __ldg(F.rel + rb + i);
    if (rb != kRelSearch) return i + static_cast<int>(-(rb + 1) - kRelShift);
    const int t = F.sn_of[rows[jj]];
    const int* trows = F.sn_rows + F.sn_rowptr[t];
    return lower_bound_rows(trows, static_cast<int>(F.sn_rowptr[t + 1] - F.sn_rowptr[t]), rows[i]);
}

extern __shared__ double madipm_smem[];

// Lx[asm_map[k]] = nzval[k]  (after a memset of Lx): the KKT values enter the panels through
// the precomputed integer map (north-star part (a), second half).
__global__ void k_assemble_panels(int64_t nnz, const int64_t* __restrict__ asm_map, const double* __restrict__ nzval,
                                  int64_t nz_stride, double* __restrict__ Lx, int64_t lx_stride) {
    nzval += blockIdx.y * nz_stride;
    Lx += blockIdx.y * lx_stride;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < nnz;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x)
        Lx[asm_map[k]] = nzval[k];
}

// inertia: signs of D.  diag_pos[k] = slot of D_kk.  out[3] = { pos, zero, neg } per instance (atomics on ints)
__global__ void k_inertia(int N, const int64_t* __restrict__ diag_pos, const double* __restrict__ Lx, int64_t lx_stride,
                          unsigned long long* __restrict__ out) {
    Lx += blockIdx.y * lx_stride;
    unsigned long long pos = 0, zero = 0, neg = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) {
        const double d = Lx[diag_pos[k]];
        if (d > 0) ++pos; else if (d < 0) ++neg; else ++zero;
    }
    for (int o = 16; o > 0; o >>= 1) {
        pos += __shfl_down_sync(0xffffffffu, pos, o);
        zero += __shfl_down_sync(0xffffffffu, zero, o);
        neg += __shfl_down_sync(0xffffffffu, neg, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[blockIdx.y * 3 + 0], pos);
        atomicAdd(&out[blockIdx.y * 3 + 1], zero);
        atomicAdd(&out[blockIdx.y * 3 + 2], neg);
    }
}

// K2b normal-equations assembly (/root/reference/src/utils.jl:266-298) as a map-driven kernel:
// Cx[c] = sum_t (A[pa[t]] * D[pk[t]]) * A[pb[t]], products and sums rounded separately and in
// the reference's summation order, so the result is bit-identical to the reference loop.
__global__ void k_assemble_normal(int64_t nnzC, const int64_t* __restrict__ prod_ptr, const int* __restrict__ pa,
                                  const int* __restrict__ pb, const int* __restrict__ pk, const double* __restrict__ Av,
                                  int64_t a_stride, const double* __restrict__ pr_diag, int64_t n_stride,
                                  double* __restrict__ Cx, int64_t c_stride) {
    Av += blockIdx.y * a_stride;
    pr_diag += blockIdx.y * n_stride;
    Cx += blockIdx.y * c_stride;
    for (int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; c < nnzC;
         c += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double acc = 0.0;
        for (int64_t t = prod_ptr[c]; t < prod_ptr[c + 1]; ++t) {
            const double dk = __ddiv_rn(1.0, pr_diag[pk[t]]);           // D .= 1.0 ./ pr_diag (normalkkt.jl:177)
            const double buf = __dmul_rn(Av[pa[t]], dk);                // buffer[j] = Jtx[c] * Dx[j]
            acc = __dadd_rn(acc, __dmul_rn(buf, Av[pb[t]]));            // Cx[c] += buffer[k] * Jtx[d]
        }
        Cx[c] = acc;
    }
}

}  // namespace madipm
