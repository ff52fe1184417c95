// Micro-benchmark (development aid): dependent-issue latency of FP64 ops, shared loads and barriers on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x, int n) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = x + i;
    __syncthreads();
    double a = x, b = x * 0.5, c = x * 0.25, d = x * 0.125;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, 1.0000001, 0.5);                     // 1 chain
    long long t1 = clock64();
    for (int i = 0; i < n; ++i) { a = fma(a, 1.0000001, 0.5); b = fma(b, 1.0000001, 0.5); c = fma(c, 1.0000001, 0.5); d = fma(d, 1.0000001, 0.5); }
    long long t2 = clock64();
    double r = x;
    for (int i = 0; i < n; ++i) { double q; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = q + 1.5; }
    long long t3 = clock64();
    double s = x;
    for (int i = 0; i < n; ++i) s = 1.0 / s + 1.5;
    long long t4 = clock64();
    int idx = threadIdx.x & 1023;
    for (int i = 0; i < n; ++i) idx = (int)sm[idx & 1023] & 1023;
    long long t5 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t6 = clock64();
    float f = (float)x;
    for (int i = 0; i < n; ++i) f = fmaf(f, 1.0000001f, 0.5f);
    long long t7 = clock64();
    for (int i = 0; i < n; ++i) __syncwarp();
    long long t8 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + r + s + idx + f;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 64);
    const int n = 4096;
    for (int threads : {32, 256, 1024}) for (int blocks : {1, 148 * 4}) {
        if (threads == 1024 && blocks > 148 * 2) blocks = 148 * 2;
        k<<<blocks, threads>>>(out, cyc, 1.25, n); cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads %4d blocks %4d | per op cycles: dfma-chain %.1f  4xdfma %.1f  rcp.approx+add %.1f  1.0/x+add %.1f  lds-chain %.1f  syncthreads %.1f  ffma %.1f syncwarp %.1f\n",
               threads, blocks, h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n, h[4] / (double)n, h[5] / (double)n, h[6] / (double)n, h[7]/(double)n);
    }
    return 0;
}
