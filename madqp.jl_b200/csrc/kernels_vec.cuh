// Fused elementwise / reduction kernels for the MPC phases (north-star part (d)) and the
// CSR products (part (c)).  All kernels take blockIdx.y = instance of the batch; every
// per-instance array is addressed base + inst * stride.
//
// Reductions are deterministic: every block writes its partials to scratch, the last block
// to arrive (atomic ticket) combines them in block order and writes the final scalars.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace madipm {

constexpr int kVecThreads = 256;
constexpr int kMaxRedBlocks = 1184;  // 148 SMs x 8 resident blocks of 256 threads

// ---------------------------------------------------------------------------------------------
// block / grid reduction helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double nanmax(double a, double b) {
    // max that propagates NaN like Julia's norm(x, Inf)
    return (a > b || a != a) ? a : b;
}

__device__ __forceinline__ double nanmin(double a, double b) { return (b < a || b != b) ? b : a; }

struct ArgMin {
    double v;
    int i;
};
// combiner of kernels.jl:248:  elem1[1] < elem2[1] ? elem1 : elem2  with elem1 the EARLIER
// element: the smaller value wins, on ties the later index wins.
__device__ __forceinline__ ArgMin argmin_combine(ArgMin a, ArgMin b) {   // a earlier than b
    return (a.v < b.v) ? a : b;
}
// The same fold as a total order (smaller value; on ties the larger index), so that partial results can be combined
// in any order: identical to the left fold for every input without NaN (a NaN direction aborts the solve anyway).
__device__ __forceinline__ ArgMin argmin_merge(ArgMin a, ArgMin b) {
    return (b.v < a.v || (b.v == a.v && b.i > a.i)) ? b : a;
}

template <int K>
struct RedVals {
    double v[K];
};

enum : int { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };

template <int OP>
__device__ __forceinline__ double red_op(double a, double b) {
    if (OP == RED_SUM) return a + b;
    if (OP == RED_MAX) return nanmax(a, b);
    return (b < a || b != b) ? b : a;           // min that propagates NaN (Julia's minimum does; the asserts of solver.jl:119-122 must see it)
}

template <int OP>
__device__ __forceinline__ double warp_red(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_op<OP>(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Reduce K values (ops given per slot) over the whole grid row (blockIdx.y fixed).
// scratch: [gridDim.y][gridDim.x][K]; counter: [gridDim.y]; out: [gridDim.y][K].
template <int K, int O0, int O1 = RED_SUM, int O2 = RED_SUM, int O3 = RED_SUM, int O4 = RED_SUM, int O5 = RED_SUM,
          int O6 = RED_SUM, int O7 = RED_SUM>
__device__ void grid_reduce(RedVals<K> r, double* scratch, unsigned int* counter, double* out) {
    __shared__ double sm[8][kVecThreads / 32];
    __shared__ bool is_last;
    constexpr int ops[8] = {O0, O1, O2, O3, O4, O5, O6, O7};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = r.v[k];
        if (ops[k] == RED_SUM) v = warp_red<RED_SUM>(v);
        else if (ops[k] == RED_MAX) v = warp_red<RED_MAX>(v);
        else v = warp_red<RED_MIN>(v);
        if (lane == 0) sm[k][warp] = v;
    }
    __syncthreads();
    const int inst = blockIdx.y;
    double* my = scratch + (static_cast<size_t>(inst) * gridDim.x + blockIdx.x) * K;
    if (threadIdx.x < K) {
        const int k = threadIdx.x;
        double v = sm[k][0];
        for (int w = 1; w < kVecThreads / 32; ++w) {
            if (ops[k] == RED_SUM) v = v + sm[k][w];
            else if (ops[k] == RED_MAX) v = nanmax(v, sm[k][w]);
            else v = (sm[k][w] < v) ? sm[k][w] : v;
        }
        my[k] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(&counter[inst], 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // warp k combines slot k: lane l walks blocks l, l+32, ... in order, then a fixed-order tree
        if (warp < K) {
            const int k = warp;
            const double* base = scratch + static_cast<size_t>(inst) * gridDim.x * K;
            const double neutral = (ops[k] == RED_SUM) ? 0.0 : ((ops[k] == RED_MAX) ? -1.0 : __longlong_as_double(0x7ff0000000000000LL));
            double v = neutral;
            for (unsigned int b0 = lane; b0 < gridDim.x; b0 += 256) {     // eight partials per lane in flight, combined in block order
                double u[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned int b = b0 + 32 * j;
                    u[j] = b < gridDim.x ? __ldcg(base + static_cast<size_t>(b) * K + k) : neutral;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (ops[k] == RED_SUM) v = v + u[j];
                    else if (ops[k] == RED_MAX) v = nanmax(v, u[j]);
                    else v = (u[j] < v) ? u[j] : v;
                }
            }
            // lanes without a block hold the neutral element of the op (0 for sums; -1 / +inf never win
            // because the reduced quantities are >= 0 for MAX and <= +inf for MIN)
            if (ops[k] == RED_SUM) v = warp_red<RED_SUM>(v);
            else if (ops[k] == RED_MAX) v = warp_red<RED_MAX>(v);
            else v = warp_red<RED_MIN>(v);
            if (lane == 0) out[static_cast<size_t>(inst) * K + k] = v;
        }
        if (threadIdx.x == 0) counter[inst] = 0;   // ready for the next launch
    }
}

// ---------------------------------------------------------------------------------------------
// CSR products  y = alpha * op(x) + beta * y, LANES threads per row (sub-warp reduction)
// ---------------------------------------------------------------------------------------------
template <int LANES>
__device__ __forceinline__ double csr_row_dot(const int* __restrict__ rp, const int* __restrict__ ci,
                                              const double* __restrict__ val, const double* __restrict__ x,
                                              int row, int sub) {
    double acc = 0.0;
    const int e = rp[row + 1];
    for (int p = rp[row] + sub; p < e; p += LANES) acc = fma(val[p], x[ci[p]], acc);
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, LANES);
    return acc;
}

// LANES threads per row, kSpmvRows rows per group in flight: the row pointers of all of them are loaded first, then
// all (column, value) pairs, then all x entries, so a thread has several independent loads outstanding instead of
// one chain rowptr -> (col, val) -> x per row.
constexpr int kSpmvRows = 4;
template <int LANES>
__global__ void __launch_bounds__(kVecThreads)
k_spmv(int rows, const int* __restrict__ rp, const int* __restrict__ ci, const double* __restrict__ val,
       int64_t val_stride, const double* __restrict__ x, int64_t x_stride, double* __restrict__ y, int64_t y_stride,
       double alpha, double beta) {
    const int inst = blockIdx.y;
    val += inst * val_stride;
    x += inst * x_stride;
    y += inst * y_stride;
    const int sub = threadIdx.x % LANES;
    constexpr int rows_per_block = kVecThreads / LANES;
    // a block owns kSpmvRows consecutive slabs of rows_per_block rows; adjacent threads keep adjacent rows in each slab
    for (int base = blockIdx.x * (kSpmvRows * rows_per_block); base < rows; base += gridDim.x * (kSpmvRows * rows_per_block)) {
        const int r0 = base + threadIdx.x / LANES;
        int p[kSpmvRows], e[kSpmvRows], c[kSpmvRows];
        double v[kSpmvRows], acc[kSpmvRows];
#pragma unroll
        for (int j = 0; j < kSpmvRows; ++j) {
            const int r = r0 + j * rows_per_block;
            p[j] = r < rows ? __ldg(rp + r) + sub : 0;
            e[j] = r < rows ? __ldg(rp + r + 1) : 0;
        }
#pragma unroll
        for (int j = 0; j < kSpmvRows; ++j) {
            const bool ok = p[j] < e[j];
            c[j] = ok ? __ldg(ci + p[j]) : 0;
            v[j] = ok ? __ldg(val + p[j]) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < kSpmvRows; ++j) acc[j] = p[j] < e[j] ? v[j] * __ldg(x + c[j]) : 0.0;
#pragma unroll
        for (int j = 0; j < kSpmvRows; ++j)
            for (int q = p[j] + LANES; q < e[j]; q += LANES) acc[j] = fma(__ldg(val + q), __ldg(x + __ldg(ci + q)), acc[j]);
#pragma unroll
        for (int j = 0; j < kSpmvRows; ++j) {
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o, LANES);
            const int r = r0 + j * rows_per_block;
            if (sub == 0 && r < rows) y[r] = (beta == 0.0) ? alpha * acc[j] : alpha * acc[j] + beta * y[r];
        }
    }
}

// gather:  dst[k] = src[map[k]]  (values of A / A' / Q from the COO value arrays)
__global__ void k_gather(int64_t n, const int* __restrict__ map, const double* __restrict__ src, int64_t src_stride,
                         double* __restrict__ dst, int64_t dst_stride) {
    const int inst = blockIdx.y;
    src += inst * src_stride;
    dst += inst * dst_stride;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[k] = src[map[k]];
}

__global__ void k_fill(int64_t n, double* __restrict__ dst, int64_t stride, double v) {
    dst += blockIdx.y * stride;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[k] = v;
}

__global__ void k_copy(int64_t n, double* __restrict__ dst, int64_t dstride, const double* __restrict__ src,
                       int64_t sstride) {
    dst += blockIdx.y * dstride;
    src += blockIdx.y * sstride;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[k] = src[k];
}

__global__ void k_axpy(int64_t n, double* __restrict__ dst, int64_t dstride, const double* __restrict__ src,
                       int64_t sstride, double alpha) {
    dst += blockIdx.y * dstride;
    src += blockIdx.y * sstride;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x)
        dst[k] += alpha * src[k];
}

__global__ void __launch_bounds__(kVecThreads)
k_norm_inf(int64_t n, const double* __restrict__ v, int64_t stride, double* scratch, unsigned int* counter, double* out) {
    v += blockIdx.y * stride;
    RedVals<1> r;
    r.v[0] = 0.0;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x)
        r.v[0] = nanmax(r.v[0], fabs(v[k]));
    grid_reduce<1, RED_MAX>(r, scratch, counter, out);
}

// ---------------------------------------------------------------------------------------------
// The per-instance view of the solver state handed to the phase kernels.
// ---------------------------------------------------------------------------------------------
struct VecView {
    int n, m, nlb, nub;
    const int* ind_lb;
    const int* ind_ub;
    const int* lb_of;   // [n] position in ind_lb or -1
    const int* ub_of;
    double *x, *xl, *xu, *zl, *zu, *f, *jacl, *cost, *y, *c, *rhs;
    double *reg, *pr_diag, *du_diag, *l_diag, *l_lower, *u_diag, *u_lower;
    double *corr_lb, *corr_ub;
    double* scaling;                // K2.5 scaling_factor [n] (unused otherwise)
    double dsg;                     // sign convention of l_diag / u_diag: +1 (K2, normal: xl - x, x - xu), -1 (K2.5: x - xl, xu - x)
    int64_t sn, sm, slb, sub, sL;   // strides (doubles) between instances
};

__device__ __forceinline__ VecView shift(VecView v, int inst) {
    v.x += inst * v.sn; v.xl += inst * v.sn; v.xu += inst * v.sn; v.zl += inst * v.sn; v.zu += inst * v.sn;
    v.f += inst * v.sn; v.jacl += inst * v.sn; v.cost += inst * v.sn;
    v.y += inst * v.sm; v.c += inst * v.sm; v.rhs += inst * v.sm;
    v.reg += inst * v.sn; v.pr_diag += inst * v.sn; v.du_diag += inst * v.sm;
    v.l_diag += inst * v.slb; v.l_lower += inst * v.slb; v.u_diag += inst * v.sub; v.u_lower += inst * v.sub;
    v.corr_lb += inst * v.slb; v.corr_ub += inst * v.sub;
    v.scaling += inst * v.sn;
    return v;
}

#define MADIPM_GRID_STRIDE(i, n) \
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += gridDim.x * blockDim.x)

// ---- K1 set_aug_diagonal_reg!  (/root/reference/src/kernels.jl:128-146), one pass.
// pr_diag[i] = (reg - zl/l_diag) - zu/u_diag in the reference's order of operations.
__global__ void k_set_aug_diagonal_reg(VecView V, double del_w, double del_c) {
    V = shift(V, blockIdx.y);
    const int top = max(V.n, V.m);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.n) {
            double pr = del_w;
            V.reg[i] = del_w;
            const int kl = V.lb_of[i], ku = V.ub_of[i];
            if (kl >= 0) {
                double ld = V.xl[i] - V.x[i];
                double ll = V.zl[i];
                V.l_diag[kl] = ld;
                V.l_lower[kl] = ll;
                pr -= ll / ld;
            }
            if (ku >= 0) {
                double ud = V.x[i] - V.xu[i];
                double ul = V.zu[i];
                V.u_diag[ku] = ud;
                V.u_lower[ku] = ul;
                pr -= ul / ud;
            }
            V.pr_diag[i] = pr;
        }
        if (i < V.m) V.du_diag[i] = del_c;
    }
}

// ---- K1 for the K2.5 system: set_aug_diagonal_reg!(::ScaledSparseKKTSystem) (/root/reference/src/kernels.jl:149-165)
// fused with MadNLP._set_aug_diagonal! [MadNLP 0.8.7, un-vendored; RECALLED, algebra forced by the symmetric scaling
// S = diag(sqrt(x - xl) * sqrt(xu - x)) of the K2 system]:  l_diag = x - xl, u_diag = xu - x (positive),
// scaling_factor = sqrt(l_diag) [* sqrt(u_diag)],  pr_diag = xlzu + xuzl + reg * scaling_factor^2  with
// xlzu = zu * (l_diag or 1), xuzl = zl * (u_diag or 1).
__global__ void k_set_aug_diagonal_scaled(VecView V, double del_w, double del_c) {
    V = shift(V, blockIdx.y);
    const int top = max(V.n, V.m);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.n) {
            V.reg[i] = del_w;
            const int kl = V.lb_of[i], ku = V.ub_of[i];
            double sf = 1.0, xlzu = 0.0, xuzl = 0.0, ld = 1.0, ud = 1.0;
            if (kl >= 0) {
                ld = V.x[i] - V.xl[i];
                const double ll = V.zl[i];
                V.l_diag[kl] = ld;
                V.l_lower[kl] = ll;
                xuzl = ll;
                sf = __dmul_rn(sf, sqrt(ld));
            }
            if (ku >= 0) {
                ud = V.xu[i] - V.x[i];
                const double ul = V.zu[i];
                V.u_diag[ku] = ud;
                V.u_lower[ku] = ul;
                xlzu = ul;
                xuzl = __dmul_rn(xuzl, ud);
                sf = __dmul_rn(sf, sqrt(ud));
            }
            if (kl >= 0) xlzu = __dmul_rn(xlzu, ld);
            V.scaling[i] = sf;
            V.pr_diag[i] = __dadd_rn(__dadd_rn(xlzu, xuzl), __dmul_rn(del_w, __dmul_rn(sf, sf)));
        }
        if (i < V.m) V.du_diag[i] = del_c;
    }
}

// ---- K9 right-hand sides (/root/reference/src/kernels.jl:1-61).  p = [px | py | pzl | pzu].
// mode 0: initial primal, 1: initial dual, 2: predictive, 3: correction (mu per instance).
__global__ void k_set_rhs(VecView V, double* __restrict__ p, int mode, const double* __restrict__ mu_dev) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    p += inst * V.sL;
    const double mu = (mode == 3) ? mu_dev[inst] : 0.0;
    double* px = p;
    double* py = p + V.n;
    double* pzl = py + V.m;
    double* pzu = pzl + V.nlb;
    const int top = max(max(V.n, V.m), max(V.nlb, V.nub));
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.n) {
            double v = 0.0;
            if (mode == 1) v = -V.f[i];
            else if (mode >= 2) v = -V.f[i] + V.zl[i] - V.zu[i] - V.jacl[i];
            px[i] = v;
        }
        if (i < V.m) py[i] = (mode == 1) ? 0.0 : -V.c[i];
        if (i < V.nlb) {
            double v = 0.0;
            if (mode >= 2) {
                const int j = V.ind_lb[i];
                v = (V.xl[j] - V.x[j]) * V.zl[j];
                if (mode == 3) v = v + mu - V.corr_lb[i];
            }
            pzl[i] = v;
        }
        if (i < V.nub) {
            double v = 0.0;
            if (mode >= 2) {
                const int j = V.ind_ub[i];
                v = (V.xu[j] - V.x[j]) * V.zu[j];
                if (mode == 3) v = v - mu - V.corr_ub[i];
            }
            pzu[i] = v;
        }
    }
}

// ---- K5 tail: _kktmul! [MadNLP, SURVEY 8a K5] per variable (no write conflicts).
// Expects wx / wy to already hold alpha * (H vx + A'vy) + beta * wx  and  alpha * A vx + beta * wy.
__global__ void k_kktmul_tail(VecView V, double* __restrict__ w, const double* __restrict__ v, double alpha, double beta) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    w += inst * V.sL;
    v += inst * V.sL;
    double* wx = w; double* wy = w + V.n; double* wzl = wy + V.m; double* wzu = wzl + V.nlb;
    const double* vx = v; const double* vy = v + V.n; const double* vzl = vy + V.m; const double* vzu = vzl + V.nlb;
    const int top = max(V.n, V.m);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.n) {
            double acc = wx[i] + alpha * V.reg[i] * vx[i];
            const int kl = V.lb_of[i], ku = V.ub_of[i];
            if (kl >= 0) {
                acc -= alpha * vzl[kl];
                wzl[kl] = beta * wzl[kl] + alpha * (vx[i] * V.l_lower[kl] - vzl[kl] * (V.dsg * V.l_diag[kl]));
            }
            if (ku >= 0) {
                acc += alpha * vzu[ku];
                wzu[ku] = beta * wzu[ku] + alpha * (vx[i] * V.u_lower[ku] + vzu[ku] * (V.dsg * V.u_diag[ku]));
            }
            wx[i] = acc;
        }
        if (i < V.m) wy[i] += alpha * V.du_diag[i] * vy[i];
    }
}

// ---- K6 norms: out = { ||w||inf, ||p||inf, ||d||inf }  (/root/reference/src/linear_solver.jl:32-34)
__global__ void __launch_bounds__(kVecThreads)
k_norms3(int64_t L, const double* __restrict__ w, const double* __restrict__ p, const double* __restrict__ d,
         int64_t stride, double* scratch, unsigned int* counter, double* out) {
    const int inst = blockIdx.y;
    w += inst * stride; p += inst * stride; d += inst * stride;
    RedVals<3> r;
    r.v[0] = r.v[1] = r.v[2] = 0.0;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < L;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        r.v[0] = nanmax(r.v[0], fabs(w[k]));
        r.v[1] = nanmax(r.v[1], fabs(p[k]));
        r.v[2] = nanmax(r.v[2], fabs(d[k]));
    }
    grid_reduce<3, RED_MAX, RED_MAX, RED_MAX>(r, scratch, counter, out);
}

// ---- K12 ratio tests (/root/reference/src/kernels.jl:242-288): four arg-min reductions with
// init (1.0, 0).  out[8] = { a_xl, a_xu, a_zl, a_zu, i_xl, i_xu, i_zl, i_zu } (1-based, 0 none).
// tau per instance.  Two-stage, deterministic, order-preserving.
__global__ void __launch_bounds__(kVecThreads)
k_ratio_test(VecView V, const double* __restrict__ d, const double* __restrict__ tau_dev,
             double* scratch /* [inst][blocks][8] */, unsigned int* counter, double* out) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    d += inst * V.sL;
    const double tau = tau_dev[inst];
    const double* dx = d;
    const double* dzl = d + V.n + V.m;
    const double* dzu = dzl + V.nlb;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    ArgMin a[4];
    // threads start from the neutral element (+inf, 0); the reference's init (1.0, 0) is applied
    // once, as the leftmost element, by the finalising block.
#pragma unroll
    for (int q = 0; q < 4; ++q) { a[q].v = inf; a[q].i = 0; }
    // coalesced grid-stride pass: a thread folds its elements in increasing order (= the reference's left fold on its
    // subset); the partial results are merged with the order-free form of the same rule
    const int top = max(V.nlb, V.nub);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < top; i += gridDim.x * blockDim.x) {
        if (i < V.nlb) {
            const int j = V.ind_lb[i];
            const double dxi = dx[j];
            ArgMin e;
            e.i = i + 1;
            e.v = (dxi < 0.0) ? (-V.x[j] + V.xl[j]) * tau / dxi : inf;
            a[0] = argmin_combine(a[0], e);
            const double dz = dzl[i];
            e.v = (dz < 0.0) ? (-V.zl[j]) * tau / dz : inf;
            a[2] = argmin_combine(a[2], e);
        }
        if (i < V.nub) {
            const int j = V.ind_ub[i];
            const double dxi = dx[j];
            ArgMin e;
            e.i = i + 1;
            e.v = (dxi > 0.0) ? (-V.x[j] + V.xu[j]) * tau / dxi : inf;
            a[1] = argmin_combine(a[1], e);
            const double dz = dzu[i];
            const double zu = V.zu[j];
            e.v = ((dz < 0.0) && (zu + dz < 0.0)) ? (-zu) * tau / dz : inf;
            a[3] = argmin_combine(a[3], e);
        }
    }
    __shared__ double sv[4][kVecThreads / 32];
    __shared__ int si[4][kVecThreads / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            ArgMin b;
            b.v = __shfl_down_sync(0xffffffffu, a[q].v, o);
            b.i = __shfl_down_sync(0xffffffffu, a[q].i, o);
            if (lane + o < 32 && ((lane & (2 * o - 1)) == 0)) a[q] = argmin_merge(a[q], b);
        }
        if (lane == 0) { sv[q][warp] = a[q].v; si[q][warp] = a[q].i; }
    }
    __syncthreads();
    double* my = scratch + (static_cast<size_t>(inst) * gridDim.x + blockIdx.x) * 8;
    if (threadIdx.x < 4) {
        const int q = threadIdx.x;
        ArgMin r; r.v = sv[q][0]; r.i = si[q][0];
        for (int w = 1; w < kVecThreads / 32; ++w) { ArgMin b; b.v = sv[q][w]; b.i = si[q][w]; r = argmin_merge(r, b); }
        my[q] = r.v;
        my[4 + q] = static_cast<double>(r.i);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&counter[inst], 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __threadfence();
        if (warp < 4) {
            // warp q merges slot q (eight partial results per lane in flight); the reference's init element (1.0, 0) is
            // applied last, as the leftmost element
            const int q = warp;
            const double* base = scratch + static_cast<size_t>(inst) * gridDim.x * 8;
            ArgMin r; r.v = inf; r.i = 0;
            for (unsigned int b0 = lane; b0 < gridDim.x; b0 += 256) {
                double ev[8], ei[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned int b = b0 + 32 * u;
                    ev[u] = b < gridDim.x ? __ldcg(base + b * 8 + q) : inf;
                    ei[u] = b < gridDim.x ? __ldcg(base + b * 8 + 4 + q) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    ArgMin e; e.v = ev[u]; e.i = static_cast<int>(ei[u]);
                    if (e.i != 0) r = argmin_merge(r, e);   // blocks that saw no element report (+inf, 0)
                }
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                ArgMin b2;
                b2.v = __shfl_down_sync(0xffffffffu, r.v, o);
                b2.i = __shfl_down_sync(0xffffffffu, r.i, o);
                if (lane + o < 32 && ((lane & (2 * o - 1)) == 0) && b2.i != 0) r = argmin_merge(r, b2);
            }
            if (lane == 0) {
                ArgMin init; init.v = 1.0; init.i = 0;
                if (r.i != 0) init = argmin_combine(init, r);
                out[static_cast<size_t>(inst) * 8 + q] = init.v;
                out[static_cast<size_t>(inst) * 8 + 4 + q] = static_cast<double>(init.i);
            }
        }
        if (threadIdx.x == 0) counter[inst] = 0;
    }
}

// ---- K11 complementarity sums (/root/reference/src/kernels.jl:171-224) + optional K10 get_correction!
// out[4] = { sum_l(alpha), sum_u(alpha), sum_l(0), sum_u(0) }.  alpha_p/alpha_d are read from
// device memory: steps[inst*2 + {0,1}] (written by the ratio-test finaliser or uploaded).
__global__ void __launch_bounds__(kVecThreads)
k_complementarity(VecView V, const double* __restrict__ d, const double* __restrict__ steps, int write_correction,
                  double* scratch, unsigned int* counter, double* out) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    d += inst * V.sL;
    const double ap = steps ? steps[inst * 2] : 0.0, ad = steps ? steps[inst * 2 + 1] : 0.0;
    const double* dx = d;
    const double* dzl = d + V.n + V.m;
    const double* dzu = dzl + V.nlb;
    RedVals<4> r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0;
    const int top = max(V.nlb, V.nub);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.nlb) {
            const int j = V.ind_lb[i];
            const double x = V.x[j], xl = V.xl[j], z = V.zl[j], dxi = dx[j], dz = dzl[i];
            r.v[0] += ((x + ap * dxi) - xl) * (z + ad * dz);
            r.v[2] += (x - xl) * z;
            if (write_correction) V.corr_lb[i] = dxi * dz;
        }
        if (i < V.nub) {
            const int j = V.ind_ub[i];
            const double x = V.x[j], xu = V.xu[j], z = V.zu[j], dxi = dx[j], dz = dzu[i];
            r.v[1] += (xu - (x + ap * dxi)) * (z + ad * dz);
            r.v[3] += (xu - x) * z;
            if (write_correction) V.corr_ub[i] = dxi * dz;
        }
    }
    grid_reduce<4, RED_SUM, RED_SUM, RED_SUM, RED_SUM>(r, scratch, counter, out);
}

// steps[inst] = { min(a_xl, a_xu), min(a_zl, a_zu) } from a ratio-test result (device side glue)
__global__ void k_steps_from_ratio(int B, const double* __restrict__ ratio_out, double* __restrict__ steps) {
    int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst < B) {
        steps[inst * 2] = fmin(ratio_out[inst * 8 + 0], ratio_out[inst * 8 + 1]);
        steps[inst * 2 + 1] = fmin(ratio_out[inst * 8 + 2], ratio_out[inst * 8 + 3]);
    }
}

// ---- device-side control scalars of one mpc! iteration (madipm_phase_iteration): what the host does between the phases when
// they are separate calls, in the same floating-point operations.
//  stage 1, after the predictor: mu_affine, mu_curr from the complementarity sums (solver.jl:296, kernels.jl:171-224),
//  update_barrier! (kernels.jl:226-236: sigma = clamp((mu_affine / mu_curr)^3, 1e-6, 10), ^3 as x*x*x like Julia's literal
//  power), tau of the step rule (kernels.jl:307-323), and the go / no-go of the iteration: every active instance factorised
//  and its predictor residual ratio is a number (linear_solver.jl:41-43).  One block; ok_all is shared by the whole batch
//  because the retry of linear_solver.jl:8-16 is (regularisation is one pair of scalars per handle).
__device__ __forceinline__ double clamp_nan(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }   // NaN stays NaN, like clamp()
__device__ __forceinline__ double max_nan(double a, double b) { return (a != a || b != b) ? (a + b) : (a > b ? a : b); }

__global__ void k_iter_scalars1(int B, int nb, const double* __restrict__ sums, const double* __restrict__ steps_aff,
                                const double* __restrict__ norms, const int* __restrict__ fail, const int* __restrict__ active,
                                int has_ineq, double mu_min, int adaptive, double tau_par, double res_tol,
                                double* __restrict__ aff4, double* __restrict__ mu, double* __restrict__ tau, int* __restrict__ ok_all) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const double mu_aff = nb ? __ddiv_rn(__dadd_rn(sums[4 * i], sums[4 * i + 1]), static_cast<double>(nb)) : 0.0;
        const double mu_cur = nb ? __ddiv_rn(__dadd_rn(sums[4 * i + 2], sums[4 * i + 3]), static_cast<double>(nb)) : 0.0;
        aff4[4 * i] = steps_aff[2 * i];
        aff4[4 * i + 1] = steps_aff[2 * i + 1];
        aff4[4 * i + 2] = mu_aff;
        aff4[4 * i + 3] = mu_cur;
        double sigma = 1.0;
        if (has_ineq) {
            const double r = __ddiv_rn(mu_aff, mu_cur);
            sigma = clamp_nan(__dmul_rn(__dmul_rn(r, r), r), 1e-6, 10.0);
        }
        const double m = max_nan(mu_min, __dmul_rn(sigma, mu_cur));
        mu[i] = m;
        tau[i] = adaptive ? max_nan(__dsub_rn(1.0, m), tau_par) : tau_par;
        const double ratio = __ddiv_rn(norms[3 * i], fmax(1.0, norms[3 * i + 1]));
        if ((!active || active[i]) && (fail[i] != 0 || ratio != ratio || (res_tol > 0.0 && ratio > res_tol))) atomicExch(&bad, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) *ok_all = bad ? 0 : 1;
}

//  stage 2, after the corrector: the steps of update_step! (kernels.jl:290-323) and the second residual check
__global__ void k_iter_scalars2(int B, const double* __restrict__ ratio8, const double* __restrict__ norms, const int* __restrict__ active,
                                double* __restrict__ steps, int* __restrict__ ok_all) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        steps[2 * i] = fmin(ratio8[8 * i], ratio8[8 * i + 1]);
        steps[2 * i + 1] = fmin(ratio8[8 * i + 2], ratio8[8 * i + 3]);
        const double ratio = __ddiv_rn(norms[3 * i], fmax(1.0, norms[3 * i + 1]));
        if ((!active || active[i]) && ratio != ratio) atomicExch(&bad, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0 && bad) *ok_all = 0;
}

// ---- Gondzio set_extra_correction! (/root/reference/src/kernels.jl:78-126)
__global__ void k_extra_correction(VecView V, const double* __restrict__ d, const double* __restrict__ steps,
                                   double bmin, double bmax, const double* __restrict__ mu_dev) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    d += inst * V.sL;
    const double ap = steps[inst * 2], ad = steps[inst * 2 + 1];
    const double mu = mu_dev[inst];
    const double tmin = bmin * mu, tmax = bmax * mu;
    const double* dx = d;
    const double* dzl = d + V.n + V.m;
    const double* dzu = dzl + V.nlb;
    const int top = max(V.nlb, V.nub);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.nlb) {
            const int j = V.ind_lb[i];
            const double x_ = V.x[j] + ap * dx[j] - V.xl[j];
            const double z_ = V.zl[j] + ad * dzl[i];
            const double v = x_ * z_;
            const double delta = (v < tmin) ? tmin - v : ((v > tmax) ? tmax - v : 0.0);
            V.corr_lb[i] = V.corr_lb[i] - delta;
        }
        if (i < V.nub) {
            const int j = V.ind_ub[i];
            const double x_ = V.xu[j] - ap * dx[j] - V.x[j];
            const double z_ = V.zu[j] + ad * dzu[i];
            const double v = x_ * z_;
            const double delta = (v < tmin) ? tmin - v : ((v > tmax) ? tmax - v : 0.0);
            V.corr_ub[i] = V.corr_ub[i] + delta;
        }
    }
}

// ---- K14 iterate update (/root/reference/src/solver.jl:332-335) fused with adjust_boundary!
// (solver.jl:342, MadNLP [RECALLED]).  steps = { alpha_p, alpha_d }, mu per instance;
// active[inst] == 0 leaves the instance untouched.
__global__ void k_update_iterate(VecView V, const double* __restrict__ d, const double* __restrict__ steps,
                                 const double* __restrict__ mu_dev, const int* __restrict__ active, const int* __restrict__ go = nullptr) {
    const int inst = blockIdx.y;
    if (go && !*go) return;                      // madipm_phase_iteration: a breakdown anywhere leaves every iterate alone
    if (active && !active[inst]) return;
    V = shift(V, inst);
    d += inst * V.sL;
    const double ap = steps[inst * 2], ad = steps[inst * 2 + 1];
    const double eps = 2.220446049250313e-16;
    const double c1 = eps * mu_dev[inst];
    const double c2 = 1.8189894035458565e-12;   // eps^(3/4)
    const double* dx = d;
    const double* dy = d + V.n;
    const double* dzl = dy + V.m;
    const double* dzu = dzl + V.nlb;
    const int top = max(V.n, V.m);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.n) {
            const double x = V.x[i] + ap * dx[i];
            V.x[i] = x;
            const int kl = V.lb_of[i], ku = V.ub_of[i];
            if (kl >= 0) {
                V.zl[i] += ad * dzl[kl];
                const double xl = V.xl[i];
                if (x - xl < c1) V.xl[i] = xl - c2 * fmax(1.0, fabs(x));
            }
            if (ku >= 0) {
                V.zu[i] += ad * dzu[ku];
                const double xu = V.xu[i];
                if (xu - x < c1) V.xu[i] = xu + c2 * fmax(1.0, fabs(x));
            }
        }
        if (i < V.m) V.y[i] += ad * dy[i];
    }
}

// dst[ind[k]] += alpha * src[k]: the `.+=` on an index view (/root/reference/src/solver.jl:334-335); ind has no repeats
__global__ void k_axpy_at(int cnt, const int* __restrict__ ind, double* __restrict__ dst, int64_t dstride,
                          const double* __restrict__ src, int64_t sstride, double alpha) {
    dst += blockIdx.y * dstride;
    src += blockIdx.y * sstride;
    MADIPM_GRID_STRIDE(k, cnt) dst[ind[k]] += alpha * src[k];
}

// MadNLP.adjust_boundary! alone (/root/reference/src/solver.jl:342): the same arithmetic as the tail of k_update_iterate
__global__ void k_adjust_boundary(VecView V, const double* __restrict__ mu_dev) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    const double eps = 2.220446049250313e-16;
    const double c1 = eps * mu_dev[inst];
    const double c2 = 1.8189894035458565e-12;   // eps^(3/4)
    MADIPM_GRID_STRIDE(i, V.n) {
        const double x = V.x[i];
        if (V.lb_of[i] >= 0) {
            const double xl = V.xl[i];
            if (x - xl < c1) V.xl[i] = xl - c2 * fmax(1.0, fabs(x));
        }
        if (V.ub_of[i] >= 0) {
            const double xu = V.xu[i];
            if (xu - x < c1) V.xu[i] = xu + c2 * fmax(1.0, fabs(x));
        }
    }
}

// ---- problem loading on the device: MadNLP.initialize!(cb, x, xl, xu, y, rhs, ind_ineq; ...) and set_scaling!
// (called at /root/reference/src/solver.jl:131-159; un-vendored MadNLP code, restated in mpc.py prepare_host, which
// is the host reference these kernels are tested against).  Raw inputs are uploaded into the vectors they end up in.
__device__ __forceinline__ double push_inside(double x, double l, double u, double bound_push, double bound_fac) {
    // Ipopt's projection of x0 into the interior of [l, u] (MadNLP.initialize_variables!)
    const bool fl = isfinite(l), fu = isfinite(u);
    const double pl = bound_push * fmax(1.0, fabs(l)), pu = bound_push * fmax(1.0, fabs(u));
    if (fl && fu) {
        const double width = bound_fac * (u - l);
        return fmin(u - fmin(pu, width), fmax(l + fmin(pl, width), x));
    }
    if (fl) return fmax(l + pl, x);
    if (fu) return fmin(u - pu, x);
    return x;
}

// x[:nx] holds x0, xl[:nx] lvar, xu[:nx] uvar on entry; relaxed bounds and the pushed start on exit; slack part zeroed
__global__ void k_load_variables(VecView V, int nx, double tol, double bound_push, double bound_fac) {
    V = shift(V, blockIdx.y);
    MADIPM_GRID_STRIDE(i, V.n) {
        if (i < nx) {
            const double l0 = V.xl[i], u0 = V.xu[i];
            const double l = l0 - fmax(1.0, fabs(l0)) * tol, u = u0 + fmax(1.0, fabs(u0)) * tol;
            V.xl[i] = l; V.xu[i] = u;
            V.x[i] = push_inside(V.x[i], l, u, bound_push, bound_fac);
        } else {
            V.x[i] = 0.0;
        }
    }
}

// slack k of inequality row r = slack_row[k]: bounds = relaxed (lcon, ucon)[r], start = cons0[r] pushed inside.
// rhs holds lcon and c holds ucon on entry (they are rewritten by k_load_scale afterwards).
__global__ void k_load_slacks(VecView V, int nx, int n_slack, const int* __restrict__ slack_row, const double* __restrict__ cons0,
                              int64_t cons_stride, double tol, double bound_push, double bound_fac) {
    cons0 += blockIdx.y * cons_stride;
    V = shift(V, blockIdx.y);
    MADIPM_GRID_STRIDE(k, n_slack) {
        const int r = slack_row[k];
        const double l0 = V.rhs[r], u0 = V.c[r];
        const double l = l0 - fmax(1.0, fabs(l0)) * tol, u = u0 + fmax(1.0, fabs(u0)) * tol;
        V.xl[nx + k] = l; V.xu[nx + k] = u;
        V.x[nx + k] = push_inside(cons0[r], l, u, bound_push, bound_fac);
    }
}

// con_scale[r] = min(1, max_gradient / max(1, max_k |A[r,k]|))   (set_scaling!, rows of the Jacobian)
__global__ void k_con_scale(int m, const int* __restrict__ rp, const double* __restrict__ val, int64_t val_stride,
                            double max_gradient, double* __restrict__ con_scale, int64_t cs_stride) {
    val += blockIdx.y * val_stride;
    con_scale += blockIdx.y * cs_stride;
    MADIPM_GRID_STRIDE(r, m) {
        double mx = 1.0;
        for (int p = rp[r]; p < rp[r + 1]; ++p) mx = fmax(mx, fabs(val[p]));
        con_scale[r] = fmin(1.0, max_gradient / mx);
    }
}

// g = Q x + cost over the structural variables, out = max |g|  (gradient of the objective at the start, for obj_scale)
__global__ void __launch_bounds__(kVecThreads)
k_grad0_max(VecView V, int nx, const int* __restrict__ Qrp, const int* __restrict__ Qci, const double* __restrict__ Qv,
            int64_t q_stride, int has_q, double* scratch, unsigned int* counter, double* out) {
    V = shift(V, blockIdx.y);
    Qv += blockIdx.y * q_stride;
    RedVals<1> r;
    r.v[0] = 0.0;
    MADIPM_GRID_STRIDE(i, nx) {
        double g = V.cost[i];
        if (has_q)
            for (int p = Qrp[i]; p < Qrp[i + 1]; ++p) g = fma(Qv[p], V.x[Qci[p]], g);
        r.v[0] = nanmax(r.v[0], fabs(g));
    }
    grid_reduce<1, RED_MAX>(r, scratch, counter, out);
}

// apply the scalings: y = y0 / cs, rhs = (lcon == ucon ? lcon : 0) * cs, slack x / xl / xu *= cs[row], cost *= obj_scale
// (rhs holds lcon and c holds ucon on entry)
__global__ void k_load_scale(VecView V, int nx, int n_slack, const int* __restrict__ slack_row, const double* __restrict__ con_scale,
                             int64_t cs_stride, const double* __restrict__ obj_scale) {
    con_scale += blockIdx.y * cs_stride;
    const double os = obj_scale[blockIdx.y];
    V = shift(V, blockIdx.y);
    const int top = max(V.n, V.m);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.m) {
            const double cs = con_scale[i];
            const double l = V.rhs[i], u = V.c[i];
            V.y[i] = V.y[i] / cs;
            V.rhs[i] = (l == u ? l : 0.0) * cs;
        }
        if (i < nx) V.cost[i] *= os;
        else if (i < V.n) {
            V.cost[i] = 0.0;
            const double cs = con_scale[slack_row[i - nx]];
            V.x[i] *= cs; V.xl[i] *= cs; V.xu[i] *= cs;
        }
    }
}

// coordinate values: hess *= obj_scale, jac[k] *= con_scale[row of k] (structural entries only)
__global__ void k_scale_values(int64_t nnzh, double* __restrict__ hessV, int64_t h_stride, int64_t nnzj, double* __restrict__ jacV,
                               int64_t j_stride, const int* __restrict__ jac_row, const double* __restrict__ con_scale, int64_t cs_stride,
                               const double* __restrict__ obj_scale) {
    hessV += blockIdx.y * h_stride;
    jacV += blockIdx.y * j_stride;
    con_scale += blockIdx.y * cs_stride;
    const double os = obj_scale[blockIdx.y];
    const int64_t top = nnzh > nnzj ? nnzh : nnzj;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < top; k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        if (k < nnzh) hessV[k] *= os;
        if (k < nnzj) jacV[k] *= con_scale[jac_row[k]];
    }
}

// ---- K8 model evaluation (/root/reference/src/solver.jl:338-340, scripts/qp_gpu.jl:29-40):
// f = Qx + cost (slack part: cost only), c = A x - rhs (slack columns are in A with -1),
// out[2] = { cost'x, x'Qx }.  One thread per row.
__global__ void __launch_bounds__(kVecThreads)
k_eval_model(VecView V, const int* __restrict__ Qrp, const int* __restrict__ Qci, const double* __restrict__ Qv,
             int64_t q_stride, const int* __restrict__ Arp, const int* __restrict__ Aci, const double* __restrict__ Av,
             int64_t a_stride, double* scratch, unsigned int* counter, double* out) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    Qv += inst * q_stride;
    Av += inst * a_stride;
    RedVals<2> r;
    r.v[0] = r.v[1] = 0.0;
    const int top = max(V.n, V.m);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.n) {
            double v = 0.0;
            for (int p = Qrp[i]; p < Qrp[i + 1]; ++p) v = fma(Qv[p], V.x[Qci[p]], v);
            const double xi = V.x[i], ci = V.cost[i];
            V.f[i] = v + ci;
            r.v[0] += ci * xi;
            r.v[1] += v * xi;
        }
        if (i < V.m) {
            double v = 0.0;
            for (int p = Arp[i]; p < Arp[i + 1]; ++p) v = fma(Av[p], V.x[Aci[p]], v);
            V.c[i] = v - V.rhs[i];
        }
    }
    grid_reduce<2, RED_SUM, RED_SUM>(r, scratch, counter, out);
}

// ---- K7 + K15 residual phase (/root/reference/src/solver.jl:259-274):
// jacl = A'y; out[4] = { ||c||inf, ||f - zl + zu + jacl||inf, max |(x-xl) zl|, |(xu-x) zu|, ||dx||inf }
__global__ void __launch_bounds__(kVecThreads)
k_residuals(VecView V, const int* __restrict__ ATrp, const int* __restrict__ ATci, const double* __restrict__ ATv,
            int64_t at_stride, const double* __restrict__ d, double* scratch, unsigned int* counter, double* out) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    ATv += inst * at_stride;
    d += inst * V.sL;
    RedVals<4> r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0;
    const int top = max(V.n, V.m);
    MADIPM_GRID_STRIDE(i, top) {
        if (i < V.n) {
            double v = 0.0;
            for (int p = ATrp[i]; p < ATrp[i + 1]; ++p) v = fma(ATv[p], V.y[ATci[p]], v);
            V.jacl[i] = v;
            r.v[1] = nanmax(r.v[1], fabs(V.f[i] - V.zl[i] + V.zu[i] + v));
            const double x = V.x[i];
            if (V.lb_of[i] >= 0) r.v[2] = nanmax(r.v[2], fabs((x - V.xl[i]) * V.zl[i]));
            if (V.ub_of[i] >= 0) r.v[2] = nanmax(r.v[2], fabs((V.xu[i] - x) * V.zu[i]));
            r.v[3] = nanmax(r.v[3], fabs(d[i]));
        }
        if (i < V.m) r.v[0] = nanmax(r.v[0], fabs(V.c[i]));
    }
    grid_reduce<4, RED_MAX, RED_MAX, RED_MAX, RED_MAX>(r, scratch, counter, out);
}

// ---- starting point, /root/reference/src/solver.jl:35-118 ------------------------------------
// stage A (l.35-78): res = jacl(=A'y, computed before) + f; init zl, zu; reductions
// out[4] = { min(x_lr - xl_r), min(xu_r - x_ur), min(zl_r), min(zu_r) } (each with init 0)
__global__ void __launch_bounds__(kVecThreads)
k_start_a(VecView V, double* scratch, unsigned int* counter, double* out) {
    V = shift(V, blockIdx.y);
    RedVals<4> r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    MADIPM_GRID_STRIDE(i, V.n) {
        const double res = V.jacl[i] + V.f[i];
        V.jacl[i] = res;
        const double l = V.xl[i], u = V.xu[i];
        const bool fl = (l > -inf) && (l < inf), fu = (u > -inf) && (u < inf);
        double zl = V.zl[i], zu = V.zu[i];
        if (fl && fu) { zl = 0.5 * res; zu = -0.5 * res; }
        else if (fl) zl = res;
        else if (fu) zu = -res;
        V.zl[i] = zl; V.zu[i] = zu;
        const double x = V.x[i];
        if (V.lb_of[i] >= 0) { r.v[0] = nanmin(r.v[0], x - l); r.v[2] = nanmin(r.v[2], zl); }
        if (V.ub_of[i] >= 0) { r.v[1] = nanmin(r.v[1], u - x); r.v[3] = nanmin(r.v[3], zu); }
    }
    grid_reduce<4, RED_MIN, RED_MIN, RED_MIN, RED_MIN>(r, scratch, counter, out);
}

// stage B (l.80-94): shifts by (delta_x, delta_s), then the sums needed for (delta_x2, delta_s2):
// out[6] = { xl.zl - lb.zl, ub.zu - xu.zu, sum zl, sum zu, sum(xl - lb), sum(ub - xu) }
__global__ void __launch_bounds__(kVecThreads)
k_start_b(VecView V, const double* __restrict__ deltas /* [inst][2] */, double* scratch, unsigned int* counter, double* out) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    const double dxs = deltas[inst * 2], dss = deltas[inst * 2 + 1];
    RedVals<6> r;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.v[k] = 0.0;
    MADIPM_GRID_STRIDE(i, V.n) {
        const bool hl = V.lb_of[i] >= 0, hu = V.ub_of[i] >= 0;
        if (!hl && !hu) continue;
        double x = V.x[i];
        if (hl) x = x + dxs;       // xl .= xl .+ delta_x   (views of the same x, applied in turn)
        if (hu) x = x - dxs;       // xu .= xu .- delta_x
        V.x[i] = x;
        if (hl) {
            const double z = V.zl[i] + (1.0 + dss);
            V.zl[i] = z;
            const double l = V.xl[i];
            r.v[0] += x * z - l * z;
            r.v[2] += z;
            r.v[4] += x - l;
        }
        if (hu) {
            const double z = V.zu[i] + (1.0 + dss);
            V.zu[i] = z;
            const double u = V.xu[i];
            r.v[1] += u * z - x * z;
            r.v[3] += z;
            r.v[5] += u - x;
        }
    }
    grid_reduce<6, RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM>(r, scratch, counter, out);
}

// stage C (l.96-123): second shift, projection into [l, u], and the four asserts as minima:
// out[4] = { min zl_r, min zu_r, min(x_lr - xl_r), min(xu_r - x_ur) } (init +inf)
__global__ void __launch_bounds__(kVecThreads)
k_start_c(VecView V, const double* __restrict__ deltas, double kappa, double* scratch, unsigned int* counter, double* out) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    const double dx2 = deltas[inst * 2], ds2 = deltas[inst * 2 + 1];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    RedVals<4> r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = inf;
    MADIPM_GRID_STRIDE(i, V.n) {
        const bool hl = V.lb_of[i] >= 0, hu = V.ub_of[i] >= 0;
        double x = V.x[i];
        if (hl) { x += dx2; V.zl[i] += ds2; }
        if (hu) { x -= dx2; V.zu[i] += ds2; }
        const double l = V.xl[i], u = V.xu[i];
        if (x < l) x = l + fmin(kappa * fmax(1.0, l), kappa * (u - l));
        else if (u < x) x = u - fmin(kappa * fmax(1.0, u), kappa * (u - l));
        V.x[i] = x;
        if (hl) { r.v[0] = nanmin(r.v[0], V.zl[i]); r.v[2] = nanmin(r.v[2], x - l); }
        if (hu) { r.v[1] = nanmin(r.v[1], V.zu[i]); r.v[3] = nanmin(r.v[3], u - x); }
    }
    grid_reduce<4, RED_MIN, RED_MIN, RED_MIN, RED_MIN>(r, scratch, counter, out);
}

// ---- KKT solve glue ---------------------------------------------------------------------------
// K2 pre-solve: xw[k] = reduce_rhs!(w)[perm[k]]   (reduce_rhs! [MadNLP, SURVEY 8a K4'], fused with the
// symmetric permutation of the linear solver)
__global__ void k_k2_presolve(VecView V, const double* __restrict__ w, const int* __restrict__ perm, int N,
                              double* __restrict__ xw, int64_t xw_stride) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    w += inst * V.sL;
    xw += inst * xw_stride;
    const double* wzl = w + V.n + V.m;
    const double* wzu = wzl + V.nlb;
    MADIPM_GRID_STRIDE(k, N) {
        const int i = perm[k];
        double v = w[i];
        if (i < V.n) {
            const int kl = V.lb_of[i], ku = V.ub_of[i];
            if (kl >= 0) v -= wzl[kl] / V.l_diag[kl];
            if (ku >= 0) v -= wzu[ku] / V.u_diag[ku];
        }
        xw[k] = v;
    }
}

// K2 post-solve: w[i] = xw[iperm[i]] and finish_aug_solve! [MadNLP, SURVEY 8a K4']
__global__ void k_k2_postsolve(VecView V, double* __restrict__ w, const int* __restrict__ iperm, int N,
                               const double* __restrict__ xw, int64_t xw_stride) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    w += inst * V.sL;
    xw += inst * xw_stride;
    double* wzl = w + V.n + V.m;
    double* wzu = wzl + V.nlb;
    MADIPM_GRID_STRIDE(i, N) {
        const double v = xw[iperm[i]];
        w[i] = v;
        if (i < V.n) {
            const int kl = V.lb_of[i], ku = V.ub_of[i];
            if (kl >= 0) wzl[kl] = (-wzl[kl] + V.l_lower[kl] * v) / V.l_diag[kl];
            if (ku >= 0) wzu[ku] = (wzu[ku] - V.u_lower[ku] * v) / V.u_diag[ku];
        }
    }
}

// K2.5 pre-solve [MadNLP solve!(::ScaledSparseKKTSystem, w), un-vendored; RECALLED, algebra forced]: with a = x - xl,
// b = xu - x, s = scaling_factor, the reduced and scaled right-hand side is  s*wx + wzl*s/a + wzu*s/b  where s/a is
// evaluated as sqrt(b)/sqrt(a) (or 1/sqrt(a)) so that a slack going to zero never meets its own square root twice.
__global__ void k_k25_presolve(VecView V, const double* __restrict__ w, const int* __restrict__ perm, int N,
                               double* __restrict__ xw, int64_t xw_stride) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    w += inst * V.sL;
    xw += inst * xw_stride;
    const double* wzl = w + V.n + V.m;
    const double* wzu = wzl + V.nlb;
    MADIPM_GRID_STRIDE(k, N) {
        const int i = perm[k];
        double v = w[i];
        if (i < V.n) {
            const int kl = V.lb_of[i], ku = V.ub_of[i];
            v *= V.scaling[i];
            const double sa = kl >= 0 ? sqrt(V.l_diag[kl]) : 1.0;
            const double sb = ku >= 0 ? sqrt(V.u_diag[ku]) : 1.0;
            if (kl >= 0) v += wzl[kl] * sb / sa;
            if (ku >= 0) v += wzu[ku] * sa / sb;
        }
        xw[k] = v;
    }
}

// K2.5 post-solve: dx = s * xw[iperm], dy = xw[iperm]; dzl = (wzl - zl*dx) / (x - xl), dzu = (zu*dx - wzu) / (xu - x)
__global__ void k_k25_postsolve(VecView V, double* __restrict__ w, const int* __restrict__ iperm, int N,
                                const double* __restrict__ xw, int64_t xw_stride) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    w += inst * V.sL;
    xw += inst * xw_stride;
    double* wzl = w + V.n + V.m;
    double* wzu = wzl + V.nlb;
    MADIPM_GRID_STRIDE(i, N) {
        double v = xw[iperm[i]];
        if (i < V.n) {
            v *= V.scaling[i];
            const int kl = V.lb_of[i], ku = V.ub_of[i];
            if (kl >= 0) wzl[kl] = (wzl[kl] - V.l_lower[kl] * v) / V.l_diag[kl];
            if (ku >= 0) wzu[ku] = (V.u_lower[ku] * v - wzu[ku]) / V.u_diag[ku];
        }
        w[i] = v;
    }
}

// Normal equations (/root/reference/src/KKT/normalkkt.jl:182-205), stage 1:
// wx <- reduce_rhs!(wx);  r1 = wx / Sigma
__global__ void k_normal_pre(VecView V, double* __restrict__ w, double* __restrict__ r1, int64_t r1_stride) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    w += inst * V.sL;
    r1 += inst * r1_stride;
    const double* wzl = w + V.n + V.m;
    const double* wzu = wzl + V.nlb;
    MADIPM_GRID_STRIDE(i, V.n) {
        double v = w[i];
        const int kl = V.lb_of[i], ku = V.ub_of[i];
        if (kl >= 0) v -= wzl[kl] / V.l_diag[kl];
        if (ku >= 0) v -= wzu[ku] / V.u_diag[ku];
        w[i] = v;
        r1[i] = v / V.pr_diag[i];
    }
}

// stage 2: xw[k] = (A r1 - wy)[perm[k]]   (one thread per row of A)
__global__ void k_normal_rhs(VecView V, const int* __restrict__ Arp, const int* __restrict__ Aci,
                             const double* __restrict__ Av, int64_t a_stride, const double* __restrict__ w,
                             const double* __restrict__ r1, int64_t r1_stride, const int* __restrict__ perm,
                             double* __restrict__ xw, int64_t xw_stride) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    Av += inst * a_stride; w += inst * V.sL; r1 += inst * r1_stride; xw += inst * xw_stride;
    const double* wy = w + V.n;
    MADIPM_GRID_STRIDE(k, V.m) {
        const int i = perm[k];
        double v = 0.0;
        for (int p = Arp[i]; p < Arp[i + 1]; ++p) v = fma(Av[p], r1[Aci[p]], v);
        xw[k] = v - wy[i];
    }
}

// stage 3: wy = xw[iperm]
__global__ void k_normal_unperm(VecView V, double* __restrict__ w, const int* __restrict__ iperm,
                                const double* __restrict__ xw, int64_t xw_stride) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    w += inst * V.sL; xw += inst * xw_stride;
    double* wy = w + V.n;
    MADIPM_GRID_STRIDE(i, V.m) wy[i] = xw[iperm[i]];
}

// stage 4: wx = (wx - A'wy) / Sigma, then finish_aug_solve!  (one thread per row of A')
__global__ void k_normal_post(VecView V, const int* __restrict__ ATrp, const int* __restrict__ ATci,
                              const double* __restrict__ ATv, int64_t at_stride, double* __restrict__ w) {
    const int inst = blockIdx.y;
    V = shift(V, inst);
    ATv += inst * at_stride; w += inst * V.sL;
    const double* wy = w + V.n;
    double* wzl = w + V.n + V.m;
    double* wzu = wzl + V.nlb;
    MADIPM_GRID_STRIDE(i, V.n) {
        double t = 0.0;
        for (int p = ATrp[i]; p < ATrp[i + 1]; ++p) t = fma(ATv[p], wy[ATci[p]], t);
        const double v = (w[i] - t) / V.pr_diag[i];
        w[i] = v;
        const int kl = V.lb_of[i], ku = V.ub_of[i];
        if (kl >= 0) wzl[kl] = (-wzl[kl] + V.l_lower[kl] * v) / V.l_diag[kl];
        if (ku >= 0) wzu[ku] = (wzu[ku] - V.u_lower[ku] * v) / V.u_diag[ku];
    }
}

// ---- MehrotraAdaptiveStep gathers (/root/reference/src/kernels.jl:341-372), one thread per instance.
// in: ratio[8] (tau = 1), mu_full sums[2] from k_complementarity at (max_alpha_p, max_alpha_d).
__global__ void k_mehrotra_step(VecView V0, int B, const double* __restrict__ d0, const double* __restrict__ ratio,
                                const double* __restrict__ compl_out, double gamma_f, double* __restrict__ out) {
    int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= B) return;
    VecView V = shift(V0, inst);
    const double* d = d0 + inst * V.sL;
    const double* dx = d;
    const double* dzl = d + V.n + V.m;
    const double* dzu = dzl + V.nlb;
    const double* r = ratio + inst * 8;
    const double axl = r[0], axu = r[1], azl = r[2], azu = r[3];
    const int ixl = static_cast<int>(r[4]) - 1, ixu = static_cast<int>(r[5]) - 1;
    const int izl = static_cast<int>(r[6]) - 1, izu = static_cast<int>(r[7]) - 1;
    const double max_ap = fmin(axl, axu), max_ad = fmin(azl, azu);
    const double gamma_a = 1.0 / (1.0 - gamma_f);
    const int nb = V.nlb + V.nub;
    double mu_full = nb ? (compl_out[inst * 4] + compl_out[inst * 4 + 1]) / nb : 0.0;
    mu_full /= gamma_a;
    double ap = 1.0, ad = 1.0;
    if (max_ap < 1.0) {
        if (axl <= axu) {
            const int j = V.ind_lb[ixl];
            const double tmp = mu_full / (V.zl[j] + max_ad * dzl[ixl]);
            ap = (V.x[j] - V.xl[j] - tmp) / (-dx[j]);
        } else {
            const int j = V.ind_ub[ixu];
            const double tmp = mu_full / (V.zu[j] + max_ad * dzu[ixu]);
            ap = (V.xu[j] - V.x[j] - tmp) / (dx[j]);
        }
    }
    if (max_ad < 1.0) {
        if (azl <= azu) {
            const int j = V.ind_lb[izl];
            const double tmp = mu_full / (V.x[j] + max_ap * dx[j] - V.xl[j]);
            ad = -(V.zl[j] - tmp) / dzl[izl];
        } else {
            const int j = V.ind_ub[izu];
            const double tmp = mu_full / (V.xu[j] - V.x[j] - max_ap * dx[j]);
            ad = -(V.zu[j] - tmp) / dzu[izu];
        }
    }
    out[inst * 2] = fmax(ap, gamma_f * max_ap);
    out[inst * 2 + 1] = fmax(ad, gamma_f * max_ad);
}

}  // namespace madipm
