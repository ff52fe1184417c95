// scale_qp on the device: Ruiz equilibration of the constraint matrix in the infinity norm and the
// rescaling of the whole model, /root/reference/scripts/common.jl:38-100.  The reference obtains (Dr, Dc) from
// HSL.mc77(A, 0) (HSL is proprietary and un-vendored; [RECALLED]: MC77 job 0 = the simultaneous row / column
// iteration of D. Ruiz, "A scaling algorithm to equilibrate both rows and columns norms in matrices",
// RAL-TR-2001-034, default 10 sweeps without a convergence test) and then divides: As = A ./ (Dr_i Dc_j),
// Hs = H ./ (Dc_i Dc_j), c ./ Dc, (lvar, uvar, x0) .* Dc, (lcon, ucon, y0) ./ Dr.
//
// standard_form_qp (common.jl:156-288) follows further down: classification of the bounds, one prefix sum, ranked writes.
//
// HBM-bound integer/FP64 streaming work: per sweep one pass over the COO triplets (16 B per entry + gathers of Dr, Dc)
// with the row / column maxima kept by 64-bit atomicMax on the bit pattern of the (non-negative) magnitude -- a maximum
// does not depend on the order of its operands, so the result is bit-reproducible and equal to the oracle's.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/madipm_b200.h"

namespace {

thread_local std::string g_scale_error;

constexpr int kThreads = 256;

inline int blocks_for(int64_t n) { return static_cast<int>(std::min<int64_t>(std::max<int64_t>((n + kThreads - 1) / kThreads, 1), 148 * 16)); }

#define GRID_STRIDE64(k, n) \
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < (n); k += static_cast<int64_t>(gridDim.x) * blockDim.x)

__global__ void k_ruiz_fill(int64_t n, double* __restrict__ x, double v) {
    GRID_STRIDE64(k, n) x[k] = v;
}

// one sweep, part 1: magnitudes of the currently scaled entries folded into the row and column maxima
__global__ void k_ruiz_max(int64_t nnz, const int* __restrict__ I, const int* __restrict__ J, const double* __restrict__ V,
                           const double* __restrict__ Dr, const double* __restrict__ Dc,
                           unsigned long long* __restrict__ rmax, unsigned long long* __restrict__ cmax) {
    GRID_STRIDE64(k, nnz) {
        const int i = I[k] - 1, j = J[k] - 1;
        const double v = fabs(V[k]) / __dmul_rn(Dr[i], Dc[j]);
        const unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(v));
        if (b > rmax[i]) atomicMax(rmax + i, b);      // (the plain read only skips atomics that cannot win)
        if (b > cmax[j]) atomicMax(cmax + j, b);
    }
}

// part 2: D *= sqrt(max) where the row / column is not empty; err = max |1 - max|; maxima reset for the next sweep
__global__ void k_ruiz_update(int64_t n, double* __restrict__ D, unsigned long long* __restrict__ mx,
                              unsigned long long* __restrict__ err) {
    double e = 0.0;
    GRID_STRIDE64(k, n) {
        const double v = __longlong_as_double(static_cast<long long>(mx[k]));
        mx[k] = 0ull;
        if (v > 0.0) {
            D[k] = __dmul_rn(D[k], sqrt(v));
            e = fmax(e, fabs(1.0 - v));
        }
    }
    for (int o = 16; o > 0; o >>= 1) e = fmax(e, __shfl_xor_sync(0xffffffffu, e, o));
    if ((threadIdx.x & 31) == 0 && e > 0.0) atomicMax(err, static_cast<unsigned long long>(__double_as_longlong(e)));
}

// _scale_coo!, common.jl:38-44
__global__ void k_scale_coo(int64_t nnz, const int* __restrict__ I, const int* __restrict__ J, double* __restrict__ V,
                            const double* __restrict__ Dr, const double* __restrict__ Dc) {
    GRID_STRIDE64(k, nnz) V[k] = V[k] / __dmul_rn(Dr[I[k] - 1], Dc[J[k] - 1]);
}

__global__ void k_scale_vec(int64_t n, double* __restrict__ x, const double* __restrict__ D, int divide) {
    GRID_STRIDE64(k, n) x[k] = divide ? x[k] / D[k] : x[k] * D[k];
}

// ---- standard_form_qp, /root/reference/scripts/common.jl:156-288 -------------------------------------------------------
// Index and copy work over the bounds: classify every variable and row (common.jl:164-199), rank the inequality rows
// (their slack columns) and the range bounds (their new rows / variables) by an exclusive prefix sum, and write the
// appended COO entries and the new bound vectors at the ranked positions.  Both ranks ride in ONE scan over the
// n + m flags [x_1..x_n | row_1..row_m]: a 64-bit value per element, low word = "inequality row" (rank k, common.jl:214),
// high word = "range bound" (rank in ind_rng: the variables first, then the slacks in row order, common.jl:176-199).
// Three kernels (tile sums, scan of the tile sums by one block, ranks + output), all streaming; ties and order cannot
// differ from the host loops because a prefix sum of integers is exact.
constexpr int kScanItems = 4;                       // elements per thread
constexpr int kScanTile = kThreads * kScanItems;    // elements per block

struct StdIn {
    int n, m;
    const double *lvar, *uvar, *lcon, *ucon;
};

// (inequality, range) flags of element e of [variables | rows], the comparisons of common.jl:166,176-183,192-195 verbatim
__device__ __forceinline__ unsigned long long std_flags(const StdIn& in, int64_t e) {
    const double ninf = -INFINITY, pinf = INFINITY;
    if (e < in.n) {
        const double l = in.lvar[e], u = in.uvar[e];
        const bool rng = !(l == u) && (ninf < l) && (l < u) && (u < pinf);
        return rng ? (1ull << 32) : 0ull;
    }
    const double l = in.lcon[e - in.n], u = in.ucon[e - in.n];
    const bool ineq = l < u;
    const bool rng = ineq && (ninf < l) && (u < pinf);
    return (ineq ? 1ull : 0ull) | (rng ? (1ull << 32) : 0ull);
}

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* total) {
    __shared__ unsigned long long warp_sum[kThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[wid] = x;
    __syncthreads();
    unsigned long long base = 0, all = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        const unsigned long long s = warp_sum[k];
        if (k < wid) base += s;
        all += s;
    }
    __syncthreads();                                  // warp_sum is reused by the next call
    *total = all;
    return base + x - v;
}

__global__ void k_std_tile_sums(StdIn in, int64_t len, unsigned long long* __restrict__ tile_sum) {
    const int64_t e0 = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
    unsigned long long s = 0;
#pragma unroll
    for (int u = 0; u < kScanItems; ++u)
        if (e0 + u < len) s += std_flags(in, e0 + u);
    unsigned long long total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place; totals[0] = (ns | nw << 32)
__global__ void k_std_scan_tiles(int64_t ntiles, unsigned long long* __restrict__ tile_sum, unsigned long long* __restrict__ totals) {
    unsigned long long carry = 0;
    for (int64_t b0 = 0; b0 < ntiles; b0 += kThreads) {
        const int64_t b = b0 + threadIdx.x;
        const unsigned long long v = b < ntiles ? tile_sum[b] : 0ull;
        unsigned long long total;
        const unsigned long long ex = block_exclusive_scan(v, &total);
        if (b < ntiles) tile_sum[b] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) totals[0] = carry;
}

struct StdOut {
    int64_t nnzj;                                    // entries of A in front of the appended ones
    int *I, *J;                                      // [nnzj + ns + 2 nw], 1-based
    double* V;
    double *c, *lvar, *uvar, *x0;                    // [n + ns + nw]; c and x0 arrive zero-filled with the model's part copied
    double *lcon, *ucon;                             // [m + nw]
};

__global__ void k_std_emit(StdIn in, int64_t len, const unsigned long long* __restrict__ tile_base,
                           const unsigned long long* __restrict__ totals, StdOut out) {
    const int n = in.n, m = in.m;
    const int ns = static_cast<int>(totals[0] & 0xffffffffull);
    const int64_t e0 = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
    unsigned long long f[kScanItems], s = 0;
#pragma unroll
    for (int u = 0; u < kScanItems; ++u) {
        f[u] = e0 + u < len ? std_flags(in, e0 + u) : 0ull;
        s += f[u];
    }
    unsigned long long total;
    unsigned long long rank = tile_base[blockIdx.x] + block_exclusive_scan(s, &total);
    const double pinf = INFINITY;
#pragma unroll
    for (int u = 0; u < kScanItems; ++u) {
        const int64_t e = e0 + u;
        if (e < len) {
            const bool ineq = f[u] & 1ull, rng = (f[u] >> 32) != 0ull;
            const int k = static_cast<int>(rank & 0xffffffffull);        // 0-based rank among the inequality rows
            const int kr = static_cast<int>(rank >> 32);                 // 0-based rank in ind_rng
            int col = 0;                                                 // 1-based column of [x | s] this element owns
            double ub = 0.0;
            if (e < n) {
                col = static_cast<int>(e) + 1;
                ub = in.uvar[e];
                out.lvar[e] = in.lvar[e];
                out.uvar[e] = rng ? pinf : ub;                           // common.jl:257 (fixed variables are never ranges)
            } else {
                const int r = static_cast<int>(e - n);
                ub = in.ucon[r];
                if (ineq) {
                    col = n + k + 1;
                    out.lcon[r] = 0.0; out.ucon[r] = 0.0;                // common.jl:239-242
                    out.lvar[n + k] = in.lcon[r];                        // common.jl:254-257
                    out.uvar[n + k] = rng ? pinf : ub;
                    const int64_t p = out.nnzj + k;                      // A x - s = 0, common.jl:214-219
                    out.I[p] = r + 1; out.J[p] = col; out.V[p] = -1.0;
                } else {
                    out.lcon[r] = in.lcon[r]; out.ucon[r] = ub;          // common.jl:244-246
                }
            }
            if (rng) {                                                   // x + w = xu, common.jl:221-230, 249-252
                const int64_t p = out.nnzj + ns + 2 * static_cast<int64_t>(kr);
                out.I[p] = m + kr + 1; out.J[p] = col; out.V[p] = 1.0;
                out.I[p + 1] = m + kr + 1; out.J[p + 1] = n + ns + kr + 1; out.V[p + 1] = 1.0;
                out.lcon[m + kr] = ub; out.ucon[m + kr] = ub;
                out.lvar[n + ns + kr] = 0.0; out.uvar[n + ns + kr] = pinf;
            }
        }
        rank += f[u];
    }
}

struct Buffers {
    void* p[40];
    int n = 0;
    ~Buffers() { for (int i = 0; i < n; ++i) cudaFree(p[i]); }
    template <class T> cudaError_t alloc(T** out, size_t count) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(out), std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) p[n++] = *out;
        return e;
    }
};

#define SC_OK(call)                                                              \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            g_scale_error = std::string(#call) + ": " + cudaGetErrorString(e_);  \
            return MADIPM_CUDA_ERROR;                                            \
        }                                                                        \
    } while (0)

}  // namespace

extern "C" {

const char* madipm_scale_last_error(void) { return g_scale_error.c_str(); }

void madipm_default_scale_options(madipm_scale_options* opt) {
    if (!opt) return;
    opt->device = 0;
    opt->max_iter = 10;     // [RECALLED] MC77 ICNTL(7)
    opt->tol = 0.0;         // [RECALLED] MC77 CNTL(1): <= 0 means no convergence test, every sweep is done
    opt->launches = 0;
}

int32_t madipm_scale_qp(madipm_scale_options* opt, int32_t nvar, int32_t ncon,
                        int64_t nnzj, const int32_t* jac_I, const int32_t* jac_J, double* jac_V,
                        int64_t nnzh, const int32_t* hess_I, const int32_t* hess_J, double* hess_V,
                        double* c, double* lvar, double* uvar, double* lcon, double* ucon, double* x0, double* y0,
                        double* Dr_out, double* Dc_out, int32_t* sweeps_done, double* final_err) {
    g_scale_error.clear();
    if (!opt || nvar < 0 || ncon < 0 || nnzj < 0 || nnzh < 0 || opt->max_iter < 0 || (nnzj > 0 && (!jac_I || !jac_J || !jac_V)) ||
        (nnzh > 0 && (!hess_I || !hess_J || !hess_V))) {
        g_scale_error = "bad argument";
        return MADIPM_BAD_ARGUMENT;
    }
    for (int64_t k = 0; k < nnzj; ++k)
        if (jac_I[k] < 1 || jac_I[k] > ncon || jac_J[k] < 1 || jac_J[k] > nvar) { g_scale_error = "Jacobian index out of range"; return MADIPM_BAD_ARGUMENT; }
    for (int64_t k = 0; k < nnzh; ++k)
        if (hess_I[k] < 1 || hess_I[k] > nvar || hess_J[k] < 1 || hess_J[k] > nvar) { g_scale_error = "Hessian index out of range"; return MADIPM_BAD_ARGUMENT; }
    int count = 0;
    cudaError_t e0 = cudaGetDeviceCount(&count);
    if (e0 != cudaSuccess || count == 0) {
        g_scale_error = std::string("no CUDA device: ") + cudaGetErrorString(e0) + " (this library has no CPU fallback)";
        return MADIPM_CUDA_ERROR;
    }
    if (opt->device < 0 || opt->device >= count) { g_scale_error = "bad device ordinal"; return MADIPM_BAD_ARGUMENT; }
    SC_OK(cudaSetDevice(opt->device));
    cudaStream_t st;
    SC_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    Buffers B;
    int64_t launches = 0;
    int *dI = nullptr, *dJ = nullptr, *dHI = nullptr, *dHJ = nullptr;
    double *dV = nullptr, *dHV = nullptr, *Dr = nullptr, *Dc = nullptr, *vn = nullptr, *vm = nullptr;
    unsigned long long *rmax = nullptr, *cmax = nullptr, *err = nullptr;
    SC_OK(B.alloc(&dI, nnzj)); SC_OK(B.alloc(&dJ, nnzj)); SC_OK(B.alloc(&dV, nnzj));
    SC_OK(B.alloc(&dHI, nnzh)); SC_OK(B.alloc(&dHJ, nnzh)); SC_OK(B.alloc(&dHV, nnzh));
    SC_OK(B.alloc(&Dr, ncon)); SC_OK(B.alloc(&Dc, nvar));
    SC_OK(B.alloc(&rmax, ncon)); SC_OK(B.alloc(&cmax, nvar)); SC_OK(B.alloc(&err, 1));
    SC_OK(B.alloc(&vn, static_cast<size_t>(nvar) * 4)); SC_OK(B.alloc(&vm, static_cast<size_t>(ncon) * 3));
    if (nnzj) {
        SC_OK(cudaMemcpyAsync(dI, jac_I, sizeof(int) * nnzj, cudaMemcpyHostToDevice, st));
        SC_OK(cudaMemcpyAsync(dJ, jac_J, sizeof(int) * nnzj, cudaMemcpyHostToDevice, st));
        SC_OK(cudaMemcpyAsync(dV, jac_V, sizeof(double) * nnzj, cudaMemcpyHostToDevice, st));
    }
    if (nnzh) {
        SC_OK(cudaMemcpyAsync(dHI, hess_I, sizeof(int) * nnzh, cudaMemcpyHostToDevice, st));
        SC_OK(cudaMemcpyAsync(dHJ, hess_J, sizeof(int) * nnzh, cudaMemcpyHostToDevice, st));
        SC_OK(cudaMemcpyAsync(dHV, hess_V, sizeof(double) * nnzh, cudaMemcpyHostToDevice, st));
    }
    SC_OK(cudaMemsetAsync(rmax, 0, sizeof(unsigned long long) * std::max(ncon, 1), st));
    SC_OK(cudaMemsetAsync(cmax, 0, sizeof(unsigned long long) * std::max(nvar, 1), st));
    k_ruiz_fill<<<blocks_for(ncon), kThreads, 0, st>>>(ncon, Dr, 1.0);
    k_ruiz_fill<<<blocks_for(nvar), kThreads, 0, st>>>(nvar, Dc, 1.0);
    launches += 2;
    int sweeps = 0;
    double last_err = 0.0;
    for (int it = 0; it < opt->max_iter; ++it) {
        SC_OK(cudaMemsetAsync(err, 0, sizeof(unsigned long long), st));
        k_ruiz_max<<<blocks_for(nnzj), kThreads, 0, st>>>(nnzj, dI, dJ, dV, Dr, Dc, rmax, cmax);
        k_ruiz_update<<<blocks_for(ncon), kThreads, 0, st>>>(ncon, Dr, rmax, err);
        k_ruiz_update<<<blocks_for(nvar), kThreads, 0, st>>>(nvar, Dc, cmax, err);
        launches += 3;
        ++sweeps;
        if (opt->tol > 0.0 || it + 1 == opt->max_iter) {       // err = deviation of the norms BEFORE this sweep's update
            unsigned long long bits = 0;
            SC_OK(cudaMemcpyAsync(&bits, err, sizeof(bits), cudaMemcpyDeviceToHost, st));
            SC_OK(cudaStreamSynchronize(st));
            double e;
            static_assert(sizeof(e) == sizeof(bits), "");
            std::memcpy(&e, &bits, sizeof(e));
            last_err = e;
            if (opt->tol > 0.0 && e <= opt->tol) break;
        }
    }
    // rescale the model (common.jl:70-97)
    if (nnzj) { k_scale_coo<<<blocks_for(nnzj), kThreads, 0, st>>>(nnzj, dI, dJ, dV, Dr, Dc); ++launches; }
    if (nnzh) { k_scale_coo<<<blocks_for(nnzh), kThreads, 0, st>>>(nnzh, dHI, dHJ, dHV, Dc, Dc); ++launches; }
    struct Item { double* host; double* dev; int64_t len; const double* D; int divide; };
    const Item items[] = {{c, vn, nvar, Dc, 1}, {lvar, vn + nvar, nvar, Dc, 0}, {uvar, vn + 2 * static_cast<int64_t>(nvar), nvar, Dc, 0},
                          {x0, vn + 3 * static_cast<int64_t>(nvar), nvar, Dc, 0}, {lcon, vm, ncon, Dr, 1},
                          {ucon, vm + ncon, ncon, Dr, 1}, {y0, vm + 2 * static_cast<int64_t>(ncon), ncon, Dr, 1}};
    for (const Item& it : items) {
        if (!it.host || it.len == 0) continue;
        SC_OK(cudaMemcpyAsync(it.dev, it.host, sizeof(double) * it.len, cudaMemcpyHostToDevice, st));
        k_scale_vec<<<blocks_for(it.len), kThreads, 0, st>>>(it.len, it.dev, it.D, it.divide);
        ++launches;
        SC_OK(cudaMemcpyAsync(it.host, it.dev, sizeof(double) * it.len, cudaMemcpyDeviceToHost, st));
    }
    if (nnzj) SC_OK(cudaMemcpyAsync(jac_V, dV, sizeof(double) * nnzj, cudaMemcpyDeviceToHost, st));
    if (nnzh) SC_OK(cudaMemcpyAsync(hess_V, dHV, sizeof(double) * nnzh, cudaMemcpyDeviceToHost, st));
    if (Dr_out && ncon) SC_OK(cudaMemcpyAsync(Dr_out, Dr, sizeof(double) * ncon, cudaMemcpyDeviceToHost, st));
    if (Dc_out && nvar) SC_OK(cudaMemcpyAsync(Dc_out, Dc, sizeof(double) * nvar, cudaMemcpyDeviceToHost, st));
    SC_OK(cudaStreamSynchronize(st));
    SC_OK(cudaGetLastError());
    if (sweeps_done) *sweeps_done = sweeps;
    if (final_err) *final_err = last_err;
    opt->launches = launches;
    return MADIPM_OK;
}

int32_t madipm_standard_form_qp(int32_t device, int32_t nvar, int32_t ncon, int64_t nnzj, const int32_t* jac_I, const int32_t* jac_J,
                                const double* jac_V, const double* c, const double* lvar, const double* uvar, const double* lcon,
                                const double* ucon, const double* x0, const double* y0, int32_t* ns_out, int32_t* nw_out,
                                int32_t* jac_I_out, int32_t* jac_J_out, double* jac_V_out, double* c_out, double* lvar_out,
                                double* uvar_out, double* x0_out, double* lcon_out, double* ucon_out, double* y0_out,
                                int64_t* launches_out) {
    g_scale_error.clear();
    const int n = nvar, m = ncon;
    if (n < 0 || m < 0 || nnzj < 0 || !ns_out || !nw_out || (n > 0 && (!lvar || !uvar)) || (m > 0 && (!lcon || !ucon)) ||
        (nnzj > 0 && (!jac_I || !jac_J || !jac_V)) || static_cast<int64_t>(n) + m >= (1ll << 30)) {
        g_scale_error = "bad argument";
        return MADIPM_BAD_ARGUMENT;
    }
    int count = 0;
    cudaError_t e0 = cudaGetDeviceCount(&count);
    if (e0 != cudaSuccess || count == 0) {
        g_scale_error = std::string("no CUDA device: ") + cudaGetErrorString(e0) + " (this library has no CPU fallback)";
        return MADIPM_CUDA_ERROR;
    }
    if (device < 0 || device >= count) { g_scale_error = "bad device ordinal"; return MADIPM_BAD_ARGUMENT; }
    SC_OK(cudaSetDevice(device));
    cudaStream_t st;
    SC_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
    Buffers B;
    const int64_t len = static_cast<int64_t>(n) + m;                   // flags [variables | rows]
    const int64_t cap_var = 2 * len, cap_con = len + m, cap_nz = nnzj + m + 2 * len;   // worst case: every row an inequality, every bound a range
    const int64_t ntiles = std::max<int64_t>(1, (len + kScanTile - 1) / kScanTile);
    double *d_lvar = nullptr, *d_uvar = nullptr, *d_lcon = nullptr, *d_ucon = nullptr;
    double *o_c = nullptr, *o_lvar = nullptr, *o_uvar = nullptr, *o_x0 = nullptr, *o_lcon = nullptr, *o_ucon = nullptr, *o_y0 = nullptr, *o_V = nullptr;
    int *o_I = nullptr, *o_J = nullptr;
    unsigned long long *tile = nullptr, *totals = nullptr;
    SC_OK(B.alloc(&d_lvar, n)); SC_OK(B.alloc(&d_uvar, n)); SC_OK(B.alloc(&d_lcon, m)); SC_OK(B.alloc(&d_ucon, m));
    SC_OK(B.alloc(&o_c, cap_var)); SC_OK(B.alloc(&o_lvar, cap_var)); SC_OK(B.alloc(&o_uvar, cap_var)); SC_OK(B.alloc(&o_x0, cap_var));
    SC_OK(B.alloc(&o_lcon, cap_con)); SC_OK(B.alloc(&o_ucon, cap_con)); SC_OK(B.alloc(&o_y0, cap_con));
    SC_OK(B.alloc(&o_I, cap_nz)); SC_OK(B.alloc(&o_J, cap_nz)); SC_OK(B.alloc(&o_V, cap_nz));
    SC_OK(B.alloc(&tile, ntiles)); SC_OK(B.alloc(&totals, 1));
    auto up = [&](void* dst, const void* src, size_t bytes) {
        return (src && bytes) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess;
    };
    SC_OK(up(d_lvar, lvar, sizeof(double) * n)); SC_OK(up(d_uvar, uvar, sizeof(double) * n));
    SC_OK(up(d_lcon, lcon, sizeof(double) * m)); SC_OK(up(d_ucon, ucon, sizeof(double) * m));
    // [c; 0], [x0; 0], [y0; 0] (common.jl:264-265, 279-280): the model's part lands in a zero-filled buffer
    SC_OK(cudaMemsetAsync(o_c, 0, sizeof(double) * std::max<int64_t>(cap_var, 1), st));
    SC_OK(cudaMemsetAsync(o_x0, 0, sizeof(double) * std::max<int64_t>(cap_var, 1), st));
    SC_OK(cudaMemsetAsync(o_y0, 0, sizeof(double) * std::max<int64_t>(cap_con, 1), st));
    SC_OK(up(o_c, c, sizeof(double) * n)); SC_OK(up(o_x0, x0, sizeof(double) * n)); SC_OK(up(o_y0, y0, sizeof(double) * m));
    // [Ai; Bi], [Aj; Bj], [Ax; Bx] (common.jl:233): the entries of A stay in front
    SC_OK(up(o_I, jac_I, sizeof(int) * nnzj)); SC_OK(up(o_J, jac_J, sizeof(int) * nnzj)); SC_OK(up(o_V, jac_V, sizeof(double) * nnzj));
    const StdIn in{n, m, d_lvar, d_uvar, d_lcon, d_ucon};
    const StdOut out{nnzj, o_I, o_J, o_V, o_c, o_lvar, o_uvar, o_x0, o_lcon, o_ucon};
    int64_t launches = 0;
    SC_OK(cudaMemsetAsync(totals, 0, sizeof(unsigned long long), st));
    if (len > 0) {
        k_std_tile_sums<<<static_cast<unsigned>(ntiles), kThreads, 0, st>>>(in, len, tile);
        k_std_scan_tiles<<<1, kThreads, 0, st>>>(ntiles, tile, totals);
        k_std_emit<<<static_cast<unsigned>(ntiles), kThreads, 0, st>>>(in, len, tile, totals, out);
        launches = 3;
    }
    unsigned long long tot = 0;
    SC_OK(cudaMemcpyAsync(&tot, totals, sizeof(tot), cudaMemcpyDeviceToHost, st));
    SC_OK(cudaStreamSynchronize(st));
    SC_OK(cudaGetLastError());
    const int64_t ns = static_cast<int64_t>(tot & 0xffffffffull), nw = static_cast<int64_t>(tot >> 32);
    const int64_t nv2 = n + ns + nw, nc2 = m + nw, nz2 = nnzj + ns + 2 * nw;
    auto down = [&](void* dst, const void* src, size_t bytes) {
        return (dst && bytes) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess;
    };
    SC_OK(down(jac_I_out, o_I, sizeof(int) * nz2)); SC_OK(down(jac_J_out, o_J, sizeof(int) * nz2)); SC_OK(down(jac_V_out, o_V, sizeof(double) * nz2));
    SC_OK(down(c_out, o_c, sizeof(double) * nv2)); SC_OK(down(lvar_out, o_lvar, sizeof(double) * nv2));
    SC_OK(down(uvar_out, o_uvar, sizeof(double) * nv2)); SC_OK(down(x0_out, o_x0, sizeof(double) * nv2));
    SC_OK(down(lcon_out, o_lcon, sizeof(double) * nc2)); SC_OK(down(ucon_out, o_ucon, sizeof(double) * nc2));
    SC_OK(down(y0_out, o_y0, sizeof(double) * nc2));
    SC_OK(cudaStreamSynchronize(st));
    *ns_out = static_cast<int32_t>(ns);
    *nw_out = static_cast<int32_t>(nw);
    if (launches_out) *launches_out = launches;
    return MADIPM_OK;
}

}  // extern "C"
