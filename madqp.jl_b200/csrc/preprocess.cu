// scale_qp on the device: Ruiz equilibration of the constraint matrix in the infinity norm and the
// rescaling of the whole model, /root/reference/scripts/common.jl:38-100.  The reference obtains (Dr, Dc) from
// HSL.mc77(A, 0) (HSL is proprietary and un-vendored; [RECALLED]: MC77 job 0 = the simultaneous row / column
// iteration of D. Ruiz, "A scaling algorithm to equilibrate both rows and columns norms in matrices",
// RAL-TR-2001-034, default 10 sweeps without a convergence test) and then divides: As = A ./ (Dr_i Dc_j),
// Hs = H ./ (Dc_i Dc_j), c ./ Dc, (lvar, uvar, x0) .* Dc, (lcon, ucon, y0) ./ Dr.
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

struct Buffers {
    void* p[24];
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

}  // extern "C"
