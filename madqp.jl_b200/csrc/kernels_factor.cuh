// Supernodal right-looking LDL' on the GPU (north-star part (b)).
//
// Storage: supernode s owns an h x w column-major panel (ld = h) at Lx + xoff[s]; its first w
// rows are its own columns.  After factorisation the panel holds D on the diagonal of the top
// w x w block, unit-lower L below it.
//
// Schedule (symbolic.cpp step 8): every bottom subtree of the supernodal elimination tree is
// walked in postorder by ONE thread block (no inter-block synchronisation inside a subtree);
// the supernodes above the subtrees are processed level by level, one launch per level.
// A supernode pushes its Schur complement L21 D L21' into its ancestors' panels with FP64
// atomics (red.global.add.f64): siblings may update the same ancestor concurrently.
//
// Panel data are read with ld.global.cg (L2) because other blocks, or this block through
// atomics, may have changed them since the line was last brought into L1.
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
    double* Lx;                // [batch][lx_stride]
    int64_t lx_stride;
    int* fail;                 // [batch]
    int h_cap;                 // rows of shared scratch available
};

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

// Factor one supernode and push its update.  Called by all threads of the block.
// shared scratch: su[h_cap] (unscaled pivot column / D), sl[max_w], rel[h_cap]
__device__ void factor_supernode(const FactorView& F, int s, double* Lx, int* fail, double* su, double* sl, int* rel) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    double* P = Lx + F.sn_xoff[s];
    const int ld = h;

    // ---- dense LDL' of the panel, one column at a time (pivot column staged in shared memory)
    for (int j = 0; j < w; ++j) {
        __syncthreads();
        const double d = ldcg(&P[j + static_cast<int64_t>(j) * ld]);
        if (tid == 0 && !(fabs(d) > 0.0 && fabs(d) < 1.0e300)) atomicExch(fail, 1);   // zero / inf / NaN pivot
        for (int i = j + 1 + tid; i < h; i += T) {
            const double u = ldcg(&P[i + static_cast<int64_t>(j) * ld]);
            const double l = u / d;
            su[i] = u;
            if (i < w) sl[i] = l;
            P[i + static_cast<int64_t>(j) * ld] = l;
        }
        __syncthreads();
        const int nrow = h - j - 1;
        const int ncol = w - j - 1;
        for (int idx = tid; idx < nrow * ncol; idx += T) {
            const int k = j + 1 + idx / nrow;
            const int i = j + 1 + idx % nrow;
            if (i >= k) {
                double* a = &P[i + static_cast<int64_t>(k) * ld];
                *a = ldcg(a) - su[i] * sl[k];
            }
        }
    }
    __syncthreads();
    if (h == w) return;

    // ---- Schur complement into the ancestors: S(i, j) = sum_k L(i,k) d_k L(j,k), i >= j >= w
    for (int k = tid; k < w; k += T) sl[k] = ldcg(&P[k + static_cast<int64_t>(k) * ld]);   // D
    int jb = w;
    while (jb < h) {
        const int t = F.sn_of[rows[jb]];
        const int fct = F.sn_first[t], lct = F.sn_first[t + 1];
        const int* trows = F.sn_rows + F.sn_rowptr[t];
        const int ht = static_cast<int>(F.sn_rowptr[t + 1] - F.sn_rowptr[t]);
        int je = jb + lower_bound_rows(rows + jb, h - jb, lct);   // rows[jb..je) are columns of t
        __syncthreads();
        for (int i = jb + tid; i < h; i += T) rel[i] = lower_bound_rows(trows, ht, rows[i]);
        __syncthreads();
        double* Pt = Lx + F.sn_xoff[t];
        const int nrow = h - jb;
        const int ncol = je - jb;
        for (int idx = tid; idx < nrow * ncol; idx += T) {
            const int jj = jb + idx / nrow;
            const int i = jb + idx % nrow;
            if (i >= jj) {
                double acc = 0.0;
                for (int k = 0; k < w; ++k)
                    acc = fma(ldcg(&P[i + static_cast<int64_t>(k) * ld]) * sl[k], ldcg(&P[jj + static_cast<int64_t>(k) * ld]), acc);
                atomicAdd(&Pt[static_cast<int64_t>(rows[jj] - fct) * ht + rel[i]], -acc);
            }
        }
        jb = je;
    }
    __syncthreads();
}

extern __shared__ unsigned char madipm_smem[];

__device__ __forceinline__ void carve(const FactorView& F, double*& su, double*& sl, int*& rel) {
    su = reinterpret_cast<double*>(madipm_smem);
    sl = su + F.h_cap;
    rel = reinterpret_cast<int*>(sl + 256);
}

__global__ void k_factor_subtrees(FactorView F, const int* __restrict__ sub_first, const int* __restrict__ sub_last) {
    const int inst = blockIdx.y;
    double* Lx = F.Lx + inst * F.lx_stride;
    double *su, *sl; int* rel;
    carve(F, su, sl, rel);
    const int first = sub_first[blockIdx.x], last = sub_last[blockIdx.x];
    for (int s = first; s <= last; ++s) {
        factor_supernode(F, s, Lx, F.fail + inst, su, sl, rel);
        __threadfence();
    }
}

__global__ void k_factor_level(FactorView F, const int* __restrict__ level_sn) {
    const int inst = blockIdx.y;
    double* Lx = F.Lx + inst * F.lx_stride;
    double *su, *sl; int* rel;
    carve(F, su, sl, rel);
    factor_supernode(F, level_sn[blockIdx.x], Lx, F.fail + inst, su, sl, rel);
}

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

// ---- triangular solves on the permuted vector xw -----------------------------------------------
__device__ void forward_supernode(const FactorView& F, int s, const double* Lx, double* xw, double* sy) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const double* P = Lx + F.sn_xoff[s];
    const int ld = h;
    __syncthreads();
    for (int j = tid; j < w; j += T) sy[j] = ldcg(&xw[fc + j]);
    for (int j = 0; j < w; ++j) {
        __syncthreads();
        const double yj = sy[j];
        for (int i = j + 1 + tid; i < w; i += T) sy[i] -= P[i + static_cast<int64_t>(j) * ld] * yj;
    }
    __syncthreads();
    for (int i = w + tid; i < h; i += T) {
        double acc = 0.0;
        for (int j = 0; j < w; ++j) acc = fma(P[i + static_cast<int64_t>(j) * ld], sy[j], acc);
        atomicAdd(&xw[rows[i]], -acc);
    }
    for (int j = tid; j < w; j += T) xw[fc + j] = sy[j] / P[j + static_cast<int64_t>(j) * ld];
}

__device__ void backward_supernode(const FactorView& F, int s, const double* Lx, double* xw, double* st) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const double* P = Lx + F.sn_xoff[s];
    const int ld = h;
    __syncthreads();
    for (int j = warp; j < w; j += nwarp) {
        double acc = 0.0;
        for (int i = w + lane; i < h; i += 32) acc = fma(P[i + static_cast<int64_t>(j) * ld], ldcg(&xw[rows[i]]), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) st[j] = ldcg(&xw[fc + j]) - acc;
    }
    for (int j = w - 1; j > 0; --j) {
        __syncthreads();
        const double xj = st[j];
        for (int k = tid; k < j; k += T) st[k] -= P[j + static_cast<int64_t>(k) * ld] * xj;
    }
    __syncthreads();
    for (int j = tid; j < w; j += T) xw[fc + j] = st[j];
}

__global__ void k_forward_subtrees(FactorView F, double* xw, int64_t xw_stride, const int* __restrict__ sub_first,
                                   const int* __restrict__ sub_last) {
    __shared__ double sy[256];
    const int inst = blockIdx.y;
    const double* Lx = F.Lx + inst * F.lx_stride;
    xw += inst * xw_stride;
    const int first = sub_first[blockIdx.x], last = sub_last[blockIdx.x];
    for (int s = first; s <= last; ++s) {
        forward_supernode(F, s, Lx, xw, sy);
        __threadfence();
    }
}

__global__ void k_forward_level(FactorView F, double* xw, int64_t xw_stride, const int* __restrict__ level_sn) {
    __shared__ double sy[256];
    const int inst = blockIdx.y;
    forward_supernode(F, level_sn[blockIdx.x], F.Lx + inst * F.lx_stride, xw + inst * xw_stride, sy);
}

__global__ void k_backward_subtrees(FactorView F, double* xw, int64_t xw_stride, const int* __restrict__ sub_first,
                                    const int* __restrict__ sub_last) {
    __shared__ double st[256];
    const int inst = blockIdx.y;
    const double* Lx = F.Lx + inst * F.lx_stride;
    xw += inst * xw_stride;
    const int first = sub_first[blockIdx.x], last = sub_last[blockIdx.x];
    for (int s = last; s >= first; --s) {
        backward_supernode(F, s, Lx, xw, st);
        __threadfence();
    }
}

__global__ void k_backward_level(FactorView F, double* xw, int64_t xw_stride, const int* __restrict__ level_sn) {
    __shared__ double st[256];
    const int inst = blockIdx.y;
    backward_supernode(F, level_sn[blockIdx.x], F.Lx + inst * F.lx_stride, xw + inst * xw_stride, st);
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
