// Micro-benchmark (development aid): cycles of dense_ldl<G> on one panel, one block.
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
__device__ long long g_probe[8];
#ifndef PROBE_JB
#define PROBE_JB 0
#endif
#define LDL_PROBE(k) do { if (t == 0 && jb == PROBE_JB) g_probe[k] = clock64(); } while (0)
#include "../../madqp.jl_b200/csrc/kernels_tree.cuh"
using namespace madipm;
template <int G>
__global__ void kb(double* P, int h, int w, int* fail, long long* cyc, int reps) {
    extern __shared__ double2 sm2[];
    double* Ps = reinterpret_cast<double*>(sm2);
    const int t = threadIdx.x % G;
    if (G < 256 && threadIdx.x >= G) return;
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
int main() {
    int shapes[][2] = {{37, 16}, {72, 64}, {109, 47}, {73, 47}, {65, 25}, {16, 4}, {24, 8}};
    for (auto& sh : shapes) {
        int h = sh[0], w = sh[1];
        std::vector<double> A(h * w), A0;
        for (int j = 0; j < w; ++j) for (int i = 0; i < h; ++i) A[i + j * h] = (i == j) ? (j % 3 == 2 ? -4.0 - j : 4.0 + j) : (i > j ? 0.3 * std::sin(1.0 + i * 7 + j * 3) : 0.0);
        A0 = A;
        double* dP; int* fail; long long* cyc;
        cudaMalloc(&dP, A.size() * 8); cudaMalloc(&fail, 4); cudaMalloc(&cyc, 8); cudaMemset(fail, 0, 4);
        for (int G : {1024, 256, 128, 32}) {
            cudaMemcpy(dP, A0.data(), A.size() * 8, cudaMemcpyHostToDevice);
            size_t smem = (h * w + 64) * 8;
            if (G == 1024) { cudaFuncSetAttribute(kb<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kb<1024><<<1, 1024, smem>>>(dP, h, w, fail, cyc, 5); }
            else if (G == 128) { cudaFuncSetAttribute(kb<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kb<128><<<1, 256, smem>>>(dP, h, w, fail, cyc, 5); }
            else if (G == 256) { cudaFuncSetAttribute(kb<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kb<256><<<1, 256, smem>>>(dP, h, w, fail, cyc, 5); }
            else { cudaFuncSetAttribute(kb<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kb<32><<<1, 32, smem>>>(dP, h, w, fail, cyc, 5); }
            cudaError_t e = cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            long long pr[8]; cudaMemcpyFromSymbol(pr, g_probe, 64);
            printf("   probed block step: update %lld | sync %lld | diag %lld | pivotcheck %lld | sync %lld | rows %lld | sync %lld\n", pr[0] - pr[7], pr[1] - pr[0], pr[2] - pr[1], pr[3] - pr[2], pr[4] - pr[3], pr[5] - pr[4], pr[6] - pr[5]);
            std::vector<double> L(A.size()); cudaMemcpy(L.data(), dP, A.size() * 8, cudaMemcpyDeviceToHost);
            // check: A0 (lower) == L D L' on the h x w panel
            double err = 0;
            for (int j = 0; j < w; ++j) for (int i = j; i < h; ++i) {
                double s = 0;
                for (int k = 0; k <= j; ++k) { double lik = (i == k) ? 1.0 : L[i + k * h], ljk = (j == k) ? 1.0 : L[j + k * h]; s += lik * L[k + k * h] * ljk; }
                err = std::fmax(err, std::fabs(s - A0[i + j * h]));
            }
            printf("h %3d w %2d G %3d: %lld cycles (%.0f per column)  err %.2e  %s\n", h, w, G, c, c / (double)w, err, cudaGetErrorString(e));
        }
    }
    return 0;
}
