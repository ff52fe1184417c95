// Micro-benchmark (development aid): cycles of dense_ldl_rr<WMAX> against dense_ldl<G> on one panel, one block.
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../../madqp.jl_b200/csrc/kernels_tree.cuh"
#include "../../madqp.jl_b200/csrc/dense_ldl_rr.cuh"
#include "../../madqp.jl_b200/csrc/dense_ldl_blk.cuh"
using namespace madipm;

template <int WMAX>
__global__ void __launch_bounds__(256, 1) kb_rr(double* P, int h, int w, long long* cyc, int reps) {
    extern __shared__ double2 sm2[];
    double* Ps = reinterpret_cast<double*>(sm2);
    double* buf = Ps + ((h * w + 1) & ~1);
    const int t = threadIdx.x;
    long long best = 1LL << 60;
    bool bad = false;
    for (int r = 0; r < reps; ++r) {
        for (int i = t; i < h * w; i += blockDim.x) Ps[i] = P[i];
        __syncthreads();
        long long t0 = clock64();
        bad |= dense_ldl_rr<WMAX>(Ps, h, w, buf, t, 1, 1e13);
        __syncthreads();
        long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
    }
    if (threadIdx.x == 0) { cyc[0] = best; cyc[1] = bad; }
    for (int i = t; i < h * w; i += blockDim.x) P[i] = Ps[i];
}

template <int G>
__global__ void kb_old(double* P, int h, int w, int* fail, long long* cyc, int reps) {
    extern __shared__ double2 sm2[];
    double* Ps = reinterpret_cast<double*>(sm2);
    const int t = threadIdx.x % G;
    long long best = 1LL << 60;
    for (int r = 0; r < reps; ++r) {
        for (int i = t; i < h * w; i += G) Ps[i] = P[i];
        gsync<G>();
        long long t0 = clock64();
        dense_ldl<G>(Ps, h, w, fail, t);
        long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
        gsync<G>();
    }
    if (threadIdx.x == 0) cyc[0] = best;
    for (int i = t; i < h * w; i += G) P[i] = Ps[i];
}

template <int G>
__global__ void kb_blk(double* P, int h, int w, int* fail, long long* cyc, int reps) {
    extern __shared__ double2 sm2[];
    double* Ps = reinterpret_cast<double*>(sm2);
    double* cb = Ps + ((h * w + 1) & ~1);
    const int t = threadIdx.x % G;
    long long best = 1LL << 60;
    for (int r = 0; r < reps; ++r) {
        for (int i = t; i < h * w; i += G) Ps[i] = P[i];
        blk_sync<G>();
        long long t0 = clock64();
        dense_ldl_blk<G>(Ps, h, w, cb, fail, t, 1e13, 1e-13);
        blk_sync<G>();
        long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
    }
    if (threadIdx.x == 0) cyc[0] = best;
    for (int i = t; i < h * w; i += G) P[i] = Ps[i];
}

static double check(const std::vector<double>& A0, const std::vector<double>& L, int h, int w) {
    double err = 0;
    for (int j = 0; j < w; ++j) for (int i = j; i < h; ++i) {
        double s = 0;
        for (int k = 0; k <= j; ++k) { double lik = (i == k) ? 1.0 : L[i + k * h], ljk = (j == k) ? 1.0 : L[j + k * h]; s += lik * L[k + k * h] * ljk; }
        err = std::fmax(err, std::fabs(s - A0[i + j * h]));
    }
    return err;
}

int main() {
    int shapes[][2] = {{109, 47}, {73, 47}, {113, 64}, {72, 64}, {104, 44}, {76, 56}, {65, 25}, {37, 16}, {24, 8}, {16, 4}, {128, 32}, {200, 40}};
    for (auto& sh : shapes) {
        int h = sh[0], w = sh[1];
        std::vector<double> A(h * w), A0;
        for (int j = 0; j < w; ++j) for (int i = 0; i < h; ++i) A[i + j * h] = (i == j) ? (j % 3 == 2 ? -4.0 - j : 4.0 + j) : (i > j ? 0.3 * std::sin(1.0 + i * 7 + j * 3) : 0.0);
        A0 = A;
        double* dP; int* fail; long long* cyc;
        cudaMalloc(&dP, A.size() * 8); cudaMalloc(&fail, 4); cudaMalloc(&cyc, 16); cudaMemset(fail, 0, 4);
        size_t smem = (h * w + 2 * 66 + 64) * 8;
        std::vector<double> L(A.size());
        long long c[2];
        // old, 256 threads
        cudaMemcpy(dP, A0.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaFuncSetAttribute(kb_old<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kb_old<256><<<1, 256, smem>>>(dP, h, w, fail, cyc, 5);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(L.data(), dP, A.size() * 8, cudaMemcpyDeviceToHost);
        printf("h %3d w %2d old<256>   : %7lld cycles (%5.0f per column) err %.2e %s\n", h, w, c[0], c[0] / (double)w, check(A0, L, h, w), cudaGetErrorString(e));
        // blocked right-looking with DMMA trailing update, 256 and (h <= 64) 32 threads
        for (int G : {256, 128, 32}) {
            if (h > 2 * G) continue;
            cudaMemcpy(dP, A0.data(), A.size() * 8, cudaMemcpyHostToDevice);
            cudaMemset(fail, 0, 4);
            if (G == 256) { cudaFuncSetAttribute(kb_blk<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kb_blk<256><<<1, 256, smem>>>(dP, h, w, fail, cyc, 5); }
            else if (G == 128) { cudaFuncSetAttribute(kb_blk<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kb_blk<128><<<1, 128, smem>>>(dP, h, w, fail, cyc, 5); }
            else { cudaFuncSetAttribute(kb_blk<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kb_blk<32><<<1, 32, smem>>>(dP, h, w, fail, cyc, 5); }
            e = cudaDeviceSynchronize();
            int hf = 0; cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(c, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(L.data(), dP, A.size() * 8, cudaMemcpyDeviceToHost);
            printf("h %3d w %2d blk<%3d>    : %7lld cycles (%5.0f per column) err %.2e fail %d %s\n", h, w, G, c[0], c[0] / (double)w, check(A0, L, h, w), hf, cudaGetErrorString(e));
        }
        // new
        cudaMemcpy(dP, A0.data(), A.size() * 8, cudaMemcpyHostToDevice);
#define RUN(WM) do { cudaFuncSetAttribute(kb_rr<WM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kb_rr<WM><<<1, 256, smem>>>(dP, h, w, cyc, 5); } while (0)
        int wm = (w + 7) / 8 * 8;
        switch (wm) { case 8: RUN(8); break; case 16: RUN(16); break; case 24: RUN(24); break; case 32: RUN(32); break; case 40: RUN(40); break;
                      case 48: RUN(48); break; case 56: RUN(56); break; default: RUN(64); }
        e = cudaDeviceSynchronize();
        cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost); cudaMemcpy(L.data(), dP, A.size() * 8, cudaMemcpyDeviceToHost);
        printf("h %3d w %2d rr<%2d>      : %7lld cycles (%5.0f per column) err %.2e bad %lld %s\n", h, w, wm, c[0], c[0] / (double)w, check(A0, L, h, w), c[1], cudaGetErrorString(e));
        cudaFree(dP); cudaFree(fail); cudaFree(cyc);
    }
    return 0;
}
