// Large fronts: supernodes whose panel does not fit one block's shared memory are factorised by
// three grid-wide kernels per level (all big supernodes of the level batched in blockIdx.y):
//
//   k_big_diag    one block per supernode: dense LDL' of the w x w diagonal block in shared memory
//   k_big_trsm    row blocks of 64: L21 = A21 Z, Z = L11^-T D^-1 from k_big_diag (a DMMA GEMM)
//   k_big_update  64 x 64 tiles of the Schur complement S = L21 D L21' on the FP64 tensor cores
//                 (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), operands staged k-major in shared memory
//                 with a leading dimension = 4 mod 16 doubles (bank-conflict-free fragment loads),
//                 scattered into the ancestors' panels through the symbolic tables.
//
// tcgen05 has no f64 kind (ptxas rejects it), so DMMA is the FP64 tensor path on sm_100a.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_factor.cuh"
#include "kernels_tree.cuh"

namespace madipm {

constexpr int kBigTile = 64;          // output tile edge of k_big_update
constexpr int kBigGroupMinRows = 2048; // ... only while that many rows remain below the group
constexpr int kBigGroup = 8;          // pieces of a cut supernode whose updates are applied together (K up to 512)
constexpr int kBigTileLd = 68;        // 64 + 4: leading dimension = 4 mod 16 doubles

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- diagonal block -----------------------------------------------------------------------------
// Factors the w x w diagonal block and also writes Z = L11^-T D^-1 (k-major, ld = w) to zbuf so that the
// triangular solve of the rows below becomes a GEMM on the tensor cores (k_big_trsm).
// shared: [w*w doubles block][w*w doubles inverse][w doubles ucol]
__global__ void k_big_diag(FactorView F, const int* __restrict__ big_sn, double* __restrict__ zbuf, int zstride) {
    const int inst = blockIdx.y;
    const int s = big_sn[blockIdx.x];
    double* Lx = F.Lx + inst * F.lx_stride;
    const int tid = threadIdx.x, T = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int w = F.sn_first[s + 1] - F.sn_first[s];
    const int64_t ld = F.sn_rowptr[s + 1] - F.sn_rowptr[s];
    double* P = Lx + F.sn_xoff[s];
    double* Bs = madipm_smem;
    double* Xs = Bs + w * w;
    double* ucol = Xs + w * w;
    for (int idx = tid; idx < w * w; idx += T) {
        const int i = idx % w, j = idx / w;
        Bs[idx] = ldcg(&P[i + j * ld]);
    }
    __syncthreads();
    for (int j = 0; j < w; ++j) {
        const double d = Bs[j * (w + 1)];
        if (tid == 0 && bad_pivot(d, F.tiny)) atomicExch(F.fail + inst, 1);
        for (int i = j + 1 + tid; i < w; i += T) {
            const double u = Bs[i + j * w];
            ucol[i] = u;
            Bs[i + j * w] = u / d;
        }
        __syncthreads();
        for (int k = j + 1 + warp; k < w; k += nwarp) {
            const double lk = Bs[k + j * w];
            for (int i = k + lane; i < w; i += 32) Bs[i + k * w] -= ucol[i] * lk;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < w * w; idx += T) {
        const int i = idx % w, j = idx / w;
        if (i >= j) P[i + j * ld] = Bs[idx];
    }
    // X = L11^-1 (unit lower), one thread per column c: x_i = -sum_{k=c}^{i-1} L[i,k] x_k
    for (int c = tid; c < w; c += T) {
        Xs[c + c * w] = 1.0;
        for (int i = c + 1; i < w; ++i) {
            double v = 0.0;
            for (int k = c; k < i; ++k) v = fma(Bs[i + k * w], Xs[k + c * w], v);
            Xs[i + c * w] = -v;
        }
    }
    __syncthreads();
    // Z[k][n] = X[n][k] / d_n for k <= n, 0 otherwise
    double* Z = zbuf + (static_cast<int64_t>(inst) * gridDim.x + blockIdx.x) * zstride;
    for (int idx = tid; idx < w * w; idx += T) {
        const int n = idx % w, k = idx / w;
        Z[k * w + n] = (k <= n) ? Xs[n + k * w] / Bs[n * (w + 1)] : 0.0;
    }
}

// ---- L21 = A21 Z,  Z = L11^-T D^-1  (GEMM on the FP64 tensor cores) --------------------------------
// grid.x = block of 64 rows below the diagonal block, grid.y = big supernode, grid.z = instance.
// 128 threads = 4 warps, each 32 rows x 32 output columns per pass over the (up to w) output columns.
// shared: [w * 68 doubles A rows, k-major][w * 68 doubles Z chunk, k-major]
__global__ void __launch_bounds__(128)
k_big_trsm(FactorView F, const int* __restrict__ big_sn, const double* __restrict__ zbuf, int zstride) {
    const int inst = blockIdx.z;
    const int s = big_sn[blockIdx.y];
    double* Lx = F.Lx + inst * F.lx_stride;
    const int tid = threadIdx.x;
    const int w = F.sn_first[s + 1] - F.sn_first[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const int64_t ld = h;
    const int r0 = w + blockIdx.x * kBigTile;
    if (r0 >= h) return;
    const int nr = min(kBigTile, h - r0);
    double* P = Lx + F.sn_xoff[s];
    const double* Z = zbuf + (static_cast<int64_t>(inst) * gridDim.y + blockIdx.y) * zstride;
    double* As = madipm_smem;
    double* Zs = As + w * kBigTileLd;
    for (int idx = tid; idx < w * kBigTile; idx += 128) {
        const int r = idx % kBigTile, k = idx / kBigTile;
        As[k * kBigTileLd + r] = (r < nr) ? ldcg(&P[r0 + r + k * ld]) : 0.0;
    }
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    for (int n0 = 0; n0 < w; n0 += kBigTile) {           // output column chunk
        const int nn = min(kBigTile, w - n0);
        __syncthreads();
        for (int idx = tid; idx < w * kBigTile; idx += 128) {
            const int n = idx % kBigTile, k = idx / kBigTile;
            Zs[k * kBigTileLd + n] = (n < nn) ? Z[k * w + n0 + n] : 0.0;
        }
        __syncthreads();
        double acc[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        const int kmax = min(w, n0 + nn);                 // Z is upper triangular: k <= n
        __syncwarp();                                     // mma.sync.aligned below: the warp must be converged
        for (int k0 = 0; k0 < kmax; k0 += 4) {
            double af[4], bf[4];
            const int kk = k0 + fk;
            const bool kin = kk < w;
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = kin ? As[kk * kBigTileLd + wm + a * 8 + fr] : 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = kin ? Zs[kk * kBigTileLd + wn + b * 8 + fr] : 0.0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int li = wm + a * 8 + fr;
                    const int lj = wn + b * 8 + 2 * fk + e;
                    if (li < nr && lj < nn) P[r0 + li + static_cast<int64_t>(n0 + lj) * ld] = acc[a][b][e];
                }
    }
}

// ---- panel factorisation of a big supernode, one launch (w <= 64) ------------------------------------------
// grid.x = block of R = 128 - w rows below the diagonal block, grid.y = big supernode, grid.z = instance; 128 threads.
// Every block stages the w x w diagonal block AND its own R rows as one (w + R) x w panel in shared memory and runs the
// same left-looking dense LDL' as the small path (dense_ldl, kernels_tree.cuh): the diagonal block is factorised
// redundantly by every block (parallel, so it costs no time) and each thread solves its own row against it -- no
// explicit inverse, no second launch.  The rows below are written back in place; the factored diagonal block goes to
// zbuf (other blocks may still be reading the unfactored one) and k_big_store_diag copies it into the panel afterwards.
constexpr int kPanelThreads = 128;
__global__ void __launch_bounds__(kPanelThreads)
k_big_panel(FactorView F, const int* __restrict__ big_sn, double* __restrict__ zbuf, int zstride) {
    extern __shared__ double2 big_smem2[];
    double* Ps = reinterpret_cast<double*>(big_smem2);
    const int inst = blockIdx.z;
    const int s = big_sn[blockIdx.y];
    double* Lx = F.Lx + inst * F.lx_stride;
    const int t = threadIdx.x;
    const int w = F.sn_first[s + 1] - F.sn_first[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const int R = kPanelThreads - w;
    const int r0 = w + blockIdx.x * R;
    if (r0 >= h && blockIdx.x > 0) return;
    const int nr = max(0, min(R, h - r0));
    const int hp = w + nr;                                 // rows of the staged panel
    double* P = Lx + F.sn_xoff[s];
    const int64_t ld = h;
    for (int idx = t; idx < hp * w; idx += kPanelThreads) {
        const int i = idx % hp, j = idx / hp;
        Ps[idx] = ldcg(&P[(i < w ? i : r0 + i - w) + j * ld]);
    }
    __syncthreads();
    dense_ldl<kPanelThreads>(Ps, hp, w, F.fail + inst, t, F.growth, F.tiny);
    __syncthreads();
    for (int idx = t; idx < nr * w; idx += kPanelThreads) {
        const int i = idx % nr, j = idx / nr;
        P[r0 + i + j * ld] = Ps[w + i + j * hp];
    }
    if (blockIdx.x == 0) {
        double* Z = zbuf + (static_cast<int64_t>(inst) * gridDim.y + blockIdx.y) * zstride;
        for (int idx = t; idx < w * w; idx += kPanelThreads) {
            const int i = idx % w, j = idx / w;
            Z[idx] = Ps[i + j * hp];
        }
    }
}

__global__ void k_big_store_diag(FactorView F, const int* __restrict__ big_sn, const double* __restrict__ zbuf, int zstride) {
    const int inst = blockIdx.y;
    const int s = big_sn[blockIdx.x];
    const int w = F.sn_first[s + 1] - F.sn_first[s];
    const int64_t ld = F.sn_rowptr[s + 1] - F.sn_rowptr[s];
    double* P = F.Lx + inst * F.lx_stride + F.sn_xoff[s];
    const double* Z = zbuf + (static_cast<int64_t>(inst) * gridDim.x + blockIdx.x) * zstride;
    for (int idx = threadIdx.x; idx < w * w; idx += blockDim.x) {
        const int i = idx % w, j = idx / w;
        if (i >= j) P[i + j * ld] = Z[idx];
    }
}

// ---- Schur complement on the FP64 tensor cores ----------------------------------------------------
// grid.x = linear index of a lower-triangular 64 x 64 tile (ti >= tj) of the (h-w) x (h-w) update,
// grid.y = big supernode of the level, grid.z = instance.  128 threads = 4 warps, each 32 x 32.
// The operands stream through shared memory in chunks of 16 columns of the panel (k), double-buffered with
// cp.async, so a block needs 35 KB instead of the whole K extent (5 blocks per SM instead of 3) and its loads
// overlap its DMMAs.  B = D L21' is formed on the fly (D in shared memory).
// shared: 2 stages x [16 x 68 doubles A][16 x 68 doubles B] (reused by the 64 x 68 epilogue tile)
//         [w doubles D][64 int64 colbase][64 int64 relbase]
constexpr int kBigKc = 16;
constexpr int kBigStage = 2 * kBigKc * kBigTileLd;        // doubles per stage (A + B)

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(128, 5)
k_big_update(FactorView F, const int* __restrict__ big_sn, const int* __restrict__ upd_lim, const int* __restrict__ upd_nsrc,
             const int* __restrict__ upd_src, int exclusive) {
    const int inst = blockIdx.z;
    const int s = big_sn[blockIdx.y];
    double* Lx = F.Lx + inst * F.lx_stride;
    const int tid = threadIdx.x;
    const int w = F.sn_first[s + 1] - F.sn_first[s];
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const int nb = h - w;
    const int nt = (nb + kBigTile - 1) / kBigTile;
    const int lim = upd_lim[blockIdx.y];                   // >= 0: only the first lim tile columns (the later pieces of the group)
    const int nsrc = upd_nsrc[blockIdx.y];                 // sources: this supernode, or every piece of its group (it is the last)
    // decode the linear tile index t -> (ti, tj), ti >= tj
    int t = blockIdx.x, ti, tj;
    if (lim < 0) {
        if (t >= nt * (nt + 1) / 2) return;
        ti = static_cast<int>((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        tj = t - ti * (ti + 1) / 2;
    } else {                                               // column by column
        tj = 0;
        while (tj < lim && tj < nt && t >= nt - tj) { t -= nt - tj; ++tj; }
        if (tj >= lim || tj >= nt) return;
        ti = tj + t;
    }
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    double* stage = madipm_smem;
    double* Ds = stage + 2 * kBigStage;                    // [nsrc][kBigTile] (one source: [w])
    const int dsz = nsrc > 1 ? kBigGroup * kBigTile : ((w + 1) & ~1);
    int64_t* colb = reinterpret_cast<int64_t*>(Ds + dsz);
    int64_t* relb = colb + kBigTile;
    const int i0 = w + ti * kBigTile, j0 = w + tj * kBigTile;
    const int ni = min(kBigTile, h - i0), nj = min(kBigTile, h - j0);
    // source q: panel, leading dimension, width, and where the target rows (the rows below s) start in it
    __shared__ int64_t src_xoff[kBigGroup];
    __shared__ int src_ld[kBigGroup], src_w[kBigGroup], src_chunk0[kBigGroup + 1];
    if (tid == 0) {
        int c = 0;
        for (int q = 0; q < nsrc; ++q) {
            const int sq = nsrc > 1 ? upd_src[blockIdx.y * kBigGroup + q] : s;
            src_xoff[q] = F.sn_xoff[sq];
            src_ld[q] = static_cast<int>(F.sn_rowptr[sq + 1] - F.sn_rowptr[sq]);
            src_w[q] = F.sn_first[sq + 1] - F.sn_first[sq];
            src_chunk0[q] = c;
            c += (src_w[q] + kBigKc - 1) / kBigKc;
        }
        src_chunk0[nsrc] = c;
    }
    __syncthreads();
    const int nchunk = src_chunk0[nsrc];
    auto issue = [&](int c) {                              // chunk c -> stage c & 1
        double* As = stage + (c & 1) * kBigStage;
        double* Bs = As + kBigKc * kBigTileLd;
        int q = 0;
        while (c >= src_chunk0[q + 1]) ++q;
        const int k0 = (c - src_chunk0[q]) * kBigKc;
        const int64_t ld = src_ld[q];
        const int wq = src_w[q];
        const double* P = Lx + src_xoff[q] + (ld - nb - w);    // row r of s's panel is row r + (ld - h) of this one
        for (int idx = tid; idx < kBigKc * kBigTile; idx += 128) {
            const int r = idx % kBigTile, k = idx / kBigTile;
            const bool kin = k0 + k < wq;
            if (kin && r < ni) cp_async8(&As[k * kBigTileLd + r], &P[i0 + r + (k0 + k) * ld]);
            else As[k * kBigTileLd + r] = 0.0;
            if (kin && r < nj) cp_async8(&Bs[k * kBigTileLd + r], &P[j0 + r + (k0 + k) * ld]);
            else Bs[k * kBigTileLd + r] = 0.0;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    if (tid < kBigTile) {
        const int jj = j0 + tid;
        colb[tid] = (tid < nj) ? F.row_colbase[F.sn_rowptr[s] + jj] : 0;
        relb[tid] = (tid < nj) ? F.row_relbase[F.sn_rowptr[s] + jj] : 0;
    }
    for (int q = 0; q < nsrc; ++q) {
        const double* Pq = Lx + src_xoff[q];
        const int64_t ld = src_ld[q];
        for (int k = tid; k < src_w[q]; k += 128) Ds[q * kBigTile + k] = ldcg(&Pq[k + k * ld]);
    }
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;     // warp's 32 x 32 corner inside the tile
    const int fr = lane >> 2, fk = lane & 3;                   // fragment row / k (A), n / k (B)
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) {
            issue(c + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const double* As = stage + (c & 1) * kBigStage;
        const double* Bs = As + kBigKc * kBigTileLd;
        int sq = 0;
        while (c >= src_chunk0[sq + 1]) ++sq;
        const double* Dq = Ds + sq * kBigTile + (c - src_chunk0[sq]) * kBigKc;
        const int kleft = src_w[sq] - (c - src_chunk0[sq]) * kBigKc;      // valid k of this chunk (the rest is zero-filled)
        __syncwarp();                                          // mma.sync.aligned below: the warp must be converged
#pragma unroll
        for (int k0 = 0; k0 < kBigKc; k0 += 4) {
            double af[4], bf[4];
            const int kk = k0 + fk;
            const double dk = Dq[min(kk, kleft - 1)];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = As[kk * kBigTileLd + wm + a * 8 + fr];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[kk * kBigTileLd + wn + b * 8 + fr] * dk;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
        __syncthreads();                                       // this stage is free for chunk c + 2
    }
    // epilogue: transpose the tile through shared memory (reusing the operand space) so that consecutive
    // lanes own consecutive rows of one target column: coalesced RMW / RED instead of 8-row fragments.
    __syncthreads();
    double* Cs = madipm_smem;                                  // [64 columns][kBigTileLd rows]
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                Cs[(wn + b * 8 + 2 * fk + e) * kBigTileLd + wm + a * 8 + fr] = acc[a][b][e];
    __syncthreads();
    const int li = tid & 63;                                   // row inside the tile
    const int i = i0 + li;
    const int cpar = tid >> 6;                                 // this thread takes columns cpar, cpar + 2, ...
    if (li < ni) {
        for (int c0 = 0; c0 < kBigTile; c0 += 16) {            // 8 columns per thread per batch
            double* q[8];
            double old[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int lj = c0 + 2 * u + cpar;
                const int jj = j0 + lj;
                q[u] = (lj < nj && i >= jj) ? Lx + colb[lj] + target_pos(F, relb[lj], i, rows, jj) : nullptr;
            }
            if (exclusive) {                                   // single supernode in this level: no other writer
#pragma unroll
                for (int u = 0; u < 8; ++u) old[u] = q[u] ? ldcg(q[u]) : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (q[u]) *q[u] = old[u] - Cs[(c0 + 2 * u + cpar) * kBigTileLd + li];
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (q[u]) atomicAdd(q[u], -Cs[(c0 + 2 * u + cpar) * kBigTileLd + li]);
            }
        }
    }
}

// ---- triangular solves with a big supernode: one launch per height and sweep ---------------------------------
// Forward: grid.x = block of 128 rows below the diagonal block (at least one block), grid.y = big supernode,
// grid.z = instance.  Every block solves the w x w unit-lower triangle redundantly (one warp, a few microseconds: it
// is the only serial part and doing it in every block saves a launch and a round trip through global memory), then
// pushes its rows' products into xw with FP64 atomics.  The last block to finish (counter) stores D^-1 y into the
// supernode's own entries of xw -- not earlier: the other blocks read the right-hand side from there.
// shared: [tri_cap doubles triangle][256 doubles y]
__global__ void __launch_bounds__(128)
k_big_forward(FactorView F, const int* __restrict__ big_sn, double* xw, int64_t xw_stride, int tri_cap, int* __restrict__ done) {
    double* tri = madipm_smem;
    double* sy = tri + tri_cap;
    __shared__ int s_last;
    const int inst = blockIdx.z;
    const int b = blockIdx.y;
    const int s = big_sn[b];
    const double* Lx = F.Lx + inst * F.lx_stride;
    xw += inst * xw_stride;
    const int tid = threadIdx.x, lane = tid & 31;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const int64_t ld = h;
    const double* P = Lx + F.sn_xoff[s];
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    const bool staged = w * w <= tri_cap;
    for (int j = tid; j < w; j += 128) sy[j] = ldcg(&xw[fc + j]);
    if (staged)
        for (int idx = tid; idx < w * w; idx += 128) {
            const int i = idx % w, j = idx / w;
            if (i > j) tri[idx] = P[i + j * ld];
        }
    __syncthreads();
    if (tid < 32) {
        for (int j = 0; j + 1 < w; ++j) {
            const double yj = sy[j];
            if (staged) { for (int i = j + 1 + lane; i < w; i += 32) sy[i] -= tri[i + j * w] * yj; }
            else { for (int i = j + 1 + lane; i < w; i += 32) sy[i] -= P[i + j * ld] * yj; }
            __syncwarp();
        }
    }
    __syncthreads();
    const int i = w + blockIdx.x * 128 + tid;
    if (i < h) {
        double acc = 0.0;
#pragma unroll 4
        for (int j = 0; j < w; ++j) acc = fma(P[i + j * ld], sy[j], acc);
        atomicAdd(&xw[rows[i]], -acc);
    }
    __threadfence();
    __syncthreads();
    int* cnt = done + inst * gridDim.y + b;
    if (tid == 0) s_last = atomicAdd(cnt, 1) == static_cast<int>(gridDim.x) - 1;
    __syncthreads();
    if (s_last) {
        for (int j = tid; j < w; j += 128) xw[fc + j] = sy[j] / P[j + j * ld];
        if (tid == 0) *cnt = 0;
    }
}

// Backward: grid.x = block of 512 rows below the diagonal block (at least one block).  Every block adds its part of
// L21' x into tbuf (FP64 atomics); the last block to finish solves the unit-upper triangle and clears tbuf.
__global__ void __launch_bounds__(256)
k_big_backward(FactorView F, const int* __restrict__ big_sn, double* __restrict__ tbuf, int tstride, double* xw, int64_t xw_stride,
               int tri_cap, int* __restrict__ done) {
    double* tri = madipm_smem;
    double* st = tri + tri_cap;
    __shared__ double sx[512];
    __shared__ int s_last;
    const int inst = blockIdx.z;
    const int b = blockIdx.y;
    const int s = big_sn[b];
    const double* Lx = F.Lx + inst * F.lx_stride;
    xw += inst * xw_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fc = F.sn_first[s];
    const int w = F.sn_first[s + 1] - fc;
    const int h = static_cast<int>(F.sn_rowptr[s + 1] - F.sn_rowptr[s]);
    const int64_t ld = h;
    const double* P = Lx + F.sn_xoff[s];
    const int* rows = F.sn_rows + F.sn_rowptr[s];
    double* tb = tbuf + (static_cast<int64_t>(inst) * gridDim.y + b) * tstride;
    const int i0 = w + blockIdx.x * 512;
    if (i0 < h) {
        const int i1 = min(h, i0 + 512);
        for (int i = i0 + tid; i < i1; i += 256) sx[i - i0] = ldcg(&xw[rows[i]]);
        __syncthreads();
        for (int j = warp; j < w; j += 8) {
            double acc = 0.0;
            for (int i = i0 + lane; i < i1; i += 32) acc = fma(P[i + j * ld], sx[i - i0], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
            if (lane == 0) atomicAdd(&tb[j], acc);
        }
    }
    __threadfence();
    __syncthreads();
    int* cnt = done + inst * gridDim.y + b;
    if (tid == 0) s_last = atomicAdd(cnt, 1) == static_cast<int>(gridDim.x) - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const bool staged = w * w <= tri_cap;
    for (int j = tid; j < w; j += 256) { st[j] = ldcg(&xw[fc + j]) - ldcg(&tb[j]); tb[j] = 0.0; }
    if (staged)
        for (int idx = tid; idx < w * w; idx += 256) {
            const int i = idx % w, j = idx / w;
            if (i > j) tri[idx] = P[i + j * ld];
        }
    __syncthreads();
    if (tid < 32) {
        for (int j = w - 1; j > 0; --j) {
            const double xj = st[j];
            if (staged) { for (int k = lane; k < j; k += 32) st[k] -= tri[j + k * w] * xj; }
            else { for (int k = lane; k < j; k += 32) st[k] -= P[j + k * ld] * xj; }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int j = tid; j < w; j += 256) xw[fc + j] = st[j];
    if (tid == 0) *cnt = 0;
}

}  // namespace madipm
