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
// atomics (red.global.add.f64): siblings may update the same ancestor concurrently.  Where an
// update lands is tabulated at symbolic time (row_colbase / row_relbase / rel), so the scatter
// has no searches and no barriers.
//
// Panels that fit are staged in shared memory: one coalesced load, the whole column loop on
// chip, one coalesced store.  Larger panels take the in-place (global memory) path.
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
    const int64_t* row_colbase;   // indexed like sn_rows
    const int64_t* row_relbase;
    const int* rel;
    double* Lx;                // [batch][lx_stride]
    int64_t lx_stride;
    int* fail;                 // [batch]
    int panel_cap;             // doubles of panel storage in shared memory (staged path)
    int scratch_rows;          // rows the in-place path can stage (pivot column)
};

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

__device__ __forceinline__ bool bad_pivot(double d) { return !(fabs(d) > 0.0 && fabs(d) < 1.0e300); }

// position of rows[i] inside the target's row list
__device__ __forceinline__ int target_pos(const FactorView& F, int64_t rb, int i, const int* rows, int jj) {
    if (rb >= 0) return __ldg(F.rel + rb + i);
    if (rb != kRelSearch) return i + static_cast<int>(-(rb + 1) - kRelShift);
    const int t = F.sn_of[rows[jj]];
    const int* trows = F.sn_rows + F.sn_rowptr[t];
    return lower_bound_rows(trows, static_cast<int>(F.sn_rowptr[t + 1] - F.sn_rowptr[t]), rows[i]);
}

// Schur complement S(i, jj) = sum_k L(i,k) d_k L(jj,k), i >= jj >= w, pushed into the ancestors.
// Lp: the factored panel (shared or global), ld = h.  Warps own columns jj, lanes own rows i.
// excl_cols: target columns below this index are owned exclusively by the calling block right now (the
// rest of its subtree, walked sequentially), so the update is a plain read-modify-write; FP64 atomics
// (which sustain only ~4e10 updates/s chip-wide) are kept for targets other blocks may also be updating.
template <bool SHARED>
__device__ __forceinline__ void push_update(const FactorView& F, int s, double* Lx, const double* Lp, const double* dk,
                                            int w, int h, const int* rows, int excl_cols) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int64_t* colb = F.row_colbase + F.sn_rowptr[s];
    const int64_t* relb = F.row_relbase + F.sn_rowptr[s];
    for (int jj = w + warp; jj < h; jj += nwarp) {
        const int64_t cb = __ldg(colb + jj);
        const int64_t rb = __ldg(relb + jj);
        const bool excl = __ldg(rows + jj) < excl_cols;
        double* dst = Lx + cb;
        for (int i = jj + lane; i < h; i += 32) {
            double acc = 0.0;
            if (SHARED) {
#pragma unroll 4
                for (int k = 0; k < w; ++k) acc = fma(Lp[i + k * h], dk[k] * Lp[jj + k * h], acc);
            } else {
                for (int k = 0; k < w; ++k)
                    acc = fma(ldcg(&Lp[i + static_cast<int64_t>(k) * h]), dk[k] * ldcg(&Lp[jj + static_cast<int64_t>(k) * h]), acc);
            }
            double* q = dst + target_pos(F, rb, i, rows, jj);
            if (excl) *q = ldcg(q) - acc;
            else atomicAdd(q, -acc);
        }
    }
}

// ---- staged path: the panel lives in shared memory for the whole column loop -------------------
__device__ void factor_supernode_staged(const FactorView& F, int s, double* Lx, int* fail, double* Ps, double* ucol, double* dk, int excl_cols) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    double* P = Lx + F.sn_xoff[s];
    const int np = h * w;
    __syncthreads();                                   // previous supernode is done with the scratch
    for (int idx = tid; idx < np; idx += T) Ps[idx] = ldcg(P + idx);
    __syncthreads();
    for (int j = 0; j < w; ++j) {
        const double d = Ps[j * (h + 1)];
        if (tid == 0 && bad_pivot(d)) atomicExch(fail, 1);
        if (j == w - 1) {
            for (int i = j + 1 + tid; i < h; i += T) Ps[i + j * h] /= d;
            break;
        }
        for (int i = j + 1 + tid; i < h; i += T) {
            const double u = Ps[i + j * h];
            ucol[i] = u;
            Ps[i + j * h] = u / d;
        }
        __syncthreads();
        for (int k = j + 1 + warp; k < w; k += nwarp) {
            const double lk = Ps[k + j * h];
            for (int i = k + lane; i < h; i += 32) Ps[i + k * h] -= ucol[i] * lk;
        }
        __syncthreads();
    }
    __syncthreads();
    for (int idx = tid; idx < np; idx += T) P[idx] = Ps[idx];
    if (h == w) return;
    for (int k = tid; k < w; k += T) dk[k] = Ps[k * (h + 1)];
    __syncthreads();
    push_update<true>(F, s, Lx, Ps, dk, w, h, rows, excl_cols);
}

// ---- in-place path: panel stays in global memory (L2); only the pivot column is staged ----------
__device__ void factor_supernode_inplace(const FactorView& F, int s, double* Lx, int* fail, double* su, double* sl, int excl_cols) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    double* P = Lx + F.sn_xoff[s];
    const int64_t ld = h;
    for (int j = 0; j < w; ++j) {
        __syncthreads();
        const double d = ldcg(&P[j + j * ld]);
        if (tid == 0 && bad_pivot(d)) atomicExch(fail, 1);
        for (int i = j + 1 + tid; i < h; i += T) {
            const double u = ldcg(&P[i + j * ld]);
            const double l = u / d;
            su[i] = u;
            if (i < w) sl[i] = l;
            P[i + j * ld] = l;
        }
        __syncthreads();
        for (int k = j + 1 + warp; k < w; k += nwarp) {
            const double lk = sl[k];
            for (int i = k + lane; i < h; i += 32) {
                double* a = &P[i + k * ld];
                *a = ldcg(a) - su[i] * lk;
            }
        }
    }
    __syncthreads();
    if (h == w) return;
    for (int k = tid; k < w; k += T) sl[k] = ldcg(&P[k + k * ld]);   // D
    __syncthreads();
    push_update<false>(F, s, Lx, P, sl, w, h, rows, excl_cols);
}

extern __shared__ double madipm_smem[];

// shared layout: [panel_cap doubles: panel | pivot column of the in-place path]
//                [panel_cap / 2 + 64 doubles: ucol][256 doubles: dk / sl]
__device__ __forceinline__ void factor_dispatch(const FactorView& F, int s, double* Lx, int* fail, int excl_cols) {
    double* Ps = madipm_smem;
    double* ucol = Ps + F.panel_cap;
    double* dk = ucol + F.panel_cap / 2 + 64;
    const int w = F.sn_first[s + 1] - F.sn_first[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const bool fits = static_cast<int64_t>(h) * w <= F.panel_cap && (w == 1 || h <= F.panel_cap / 2 + 64);
    if (fits) factor_supernode_staged(F, s, Lx, fail, Ps, ucol, dk, excl_cols);
    else factor_supernode_inplace(F, s, Lx, fail, Ps, dk, excl_cols);
}

__global__ void k_factor_subtrees(FactorView F, const int* __restrict__ sub_first, const int* __restrict__ sub_last) {
    const int inst = blockIdx.y;
    double* Lx = F.Lx + inst * F.lx_stride;
    const int first = sub_first[blockIdx.x], last = sub_last[blockIdx.x];
    const int excl_cols = F.sn_first[last + 1];     // columns of this subtree: nobody else touches their panels now
    for (int s = first; s <= last; ++s) {
        factor_dispatch(F, s, Lx, F.fail + inst, excl_cols);
        __threadfence();
    }
}

// exclusive != 0: the level holds a single supernode, so every target is exclusively ours
__global__ void k_factor_level(FactorView F, const int* __restrict__ level_sn, int exclusive) {
    const int inst = blockIdx.y;
    factor_dispatch(F, level_sn[blockIdx.x], F.Lx + inst * F.lx_stride, F.fail + inst, exclusive ? INT32_MAX : 0);
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
// tri: w x w scratch for the diagonal block (staged with one coalesced pass), sy: w values.
__device__ void forward_supernode(const FactorView& F, int s, const double* Lx, double* xw, double* tri, double* sy, int tri_cap) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const double* P = Lx + F.sn_xoff[s];
    const int64_t ld = h;
    const bool staged = w * w <= tri_cap;
    __syncthreads();
    for (int j = tid; j < w; j += T) sy[j] = ldcg(&xw[fc + j]);
    if (staged && w > 1)
        for (int idx = tid; idx < w * w; idx += T) {
            const int i = idx % w, j = idx / w;
            if (i > j) tri[idx] = P[i + j * ld];
        }
    for (int j = 0; j + 1 < w; ++j) {
        __syncthreads();
        const double yj = sy[j];
        if (staged) { for (int i = j + 1 + tid; i < w; i += T) sy[i] -= tri[i + j * w] * yj; }
        else { for (int i = j + 1 + tid; i < w; i += T) sy[i] -= P[i + j * ld] * yj; }
    }
    __syncthreads();
    for (int i = w + tid; i < h; i += T) {
        double acc = 0.0;
        for (int j = 0; j < w; ++j) acc = fma(P[i + j * ld], sy[j], acc);
        atomicAdd(&xw[rows[i]], -acc);
    }
    for (int j = tid; j < w; j += T) xw[fc + j] = sy[j] / P[j + j * ld];
}

__device__ void backward_supernode(const FactorView& F, int s, const double* Lx, double* xw, double* tri, double* st, int tri_cap) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const double* P = Lx + F.sn_xoff[s];
    const int64_t ld = h;
    const bool staged = w * w <= tri_cap;
    __syncthreads();
    if (staged && w > 1)
        for (int idx = tid; idx < w * w; idx += T) {
            const int i = idx % w, j = idx / w;
            if (i > j) tri[idx] = P[i + j * ld];
        }
    for (int j = warp; j < w; j += nwarp) {
        double acc = 0.0;
        for (int i = w + lane; i < h; i += 32) acc = fma(P[i + j * ld], ldcg(&xw[rows[i]]), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) st[j] = ldcg(&xw[fc + j]) - acc;
    }
    for (int j = w - 1; j > 0; --j) {
        __syncthreads();
        const double xj = st[j];
        if (staged) { for (int k = tid; k < j; k += T) st[k] -= tri[j + k * w] * xj; }
        else { for (int k = tid; k < j; k += T) st[k] -= P[j + k * ld] * xj; }
    }
    __syncthreads();
    for (int j = tid; j < w; j += T) xw[fc + j] = st[j];
}

// dynamic shared memory of the solve kernels: [tri_cap doubles][256 doubles]
__global__ void k_forward_subtrees(FactorView F, double* xw, int64_t xw_stride, const int* __restrict__ sub_first,
                                   const int* __restrict__ sub_last, int tri_cap) {
    double* tri = madipm_smem;
    double* sy = tri + tri_cap;
    const int inst = blockIdx.y;
    const double* Lx = F.Lx + inst * F.lx_stride;
    xw += inst * xw_stride;
    const int first = sub_first[blockIdx.x], last = sub_last[blockIdx.x];
    for (int s = first; s <= last; ++s) {
        forward_supernode(F, s, Lx, xw, tri, sy, tri_cap);
        __threadfence();
    }
}

__global__ void k_forward_level(FactorView F, double* xw, int64_t xw_stride, const int* __restrict__ level_sn, int tri_cap) {
    double* tri = madipm_smem;
    const int inst = blockIdx.y;
    forward_supernode(F, level_sn[blockIdx.x], F.Lx + inst * F.lx_stride, xw + inst * xw_stride, tri, tri + tri_cap, tri_cap);
}

__global__ void k_backward_subtrees(FactorView F, double* xw, int64_t xw_stride, const int* __restrict__ sub_first,
                                    const int* __restrict__ sub_last, int tri_cap) {
    double* tri = madipm_smem;
    double* st = tri + tri_cap;
    const int inst = blockIdx.y;
    const double* Lx = F.Lx + inst * F.lx_stride;
    xw += inst * xw_stride;
    const int first = sub_first[blockIdx.x], last = sub_last[blockIdx.x];
    for (int s = last; s >= first; --s) {
        backward_supernode(F, s, Lx, xw, tri, st, tri_cap);
        __threadfence();
    }
}

__global__ void k_backward_level(FactorView F, double* xw, int64_t xw_stride, const int* __restrict__ level_sn, int tri_cap) {
    double* tri = madipm_smem;
    const int inst = blockIdx.y;
    backward_supernode(F, level_sn[blockIdx.x], F.Lx + inst * F.lx_stride, xw + inst * xw_stride, tri, tri + tri_cap, tri_cap);
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
