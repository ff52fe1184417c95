// Dense LDL' of one h x w front in shared memory for the task groups of k_tree (64 registers per thread): right-looking in
// blocks of 8 columns.
//   (1) panel step: the 8 columns of the block are factorised with the rows in registers (8 entries per row, at most two rows
//       per thread), one barrier per column; the pivot column is published through a 2 x 8 shared buffer;
//   (2) trailing update  A22 -= L21 (D L21')  on the FP64 tensor cores: 8 x 8 tiles of the lower triangle dealt to the warps of
//       the group, two DMMA.8x8x4 per tile (k = 8), accumulators independent from tile to tile (no dependent DMMA chains:
//       the left-looking variant of round 1 had w / 4 of them per tile).
// The blocked LEFT-looking routine (dense_ldl) reads one own-row and four broadcast shared-memory operands per four FMAs, for
// every earlier column again (w^2 h / 2 FMAs, 1.25 shared loads each): its load/store pipe is the bottleneck.  Here the FMAs
// of the update run on the tensor pipe at two shared loads per 32 FMAs of a lane.
#pragma once
#include <cuda_runtime.h>

#include "dense_ldl_rr.cuh"
#include "kernels_factor.cuh"

namespace madipm {

constexpr int kBlkNb = 8;

template <int G>
__device__ __forceinline__ void blk_sync() {
    if (G == 32) __syncwarp();
    else if (G >= 256) __syncthreads();
    else named_barrier((G == 128 ? 1 : 3) + static_cast<int>(threadIdx.x) / G, G);
}
// barrier among the first nact threads (a multiple of 32) of the group: the warps that own rows in the panel step.  Same
// auxiliary barrier ids as gsync_active in kernels_tree.cuh (7-8: the two 128-thread groups of a block, 9: a 256- or
// 1024-thread group; a 64-thread group is one or two warps).
template <int G>
__device__ __forceinline__ void blk_sync_active(int nact) {
    if (G == 32 || nact <= 32) __syncwarp();
    else if (nact >= G) blk_sync<G>();
    else named_barrier(G == 128 ? 7 + static_cast<int>(threadIdx.x) / G : 9, nact);
}

// Ps: panel (ld = h).  cb: 2 * 8 + 8 doubles of shared memory private to the group (16-byte aligned).  Requires h <= 2 * G.
// On return D on the diagonal, unit-lower L below; the strict upper triangle of the pivot block is scratch.
template <int G>
__device__ __forceinline__ void dense_ldl_blk(double* __restrict__ Ps, int h, int w, double* __restrict__ cb, int* fail, int t,
                                              double growth, double tiny) {
    constexpr int NW = G / 32;
    const int lane = t & 31, gw = t >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    double* dblk = cb + 16;                                   // D of the current block (8 doubles)
    bool bad = false;
    for (int jb = 0; jb < w; jb += kBlkNb) {
        const int bw = min(kBlkNb, w - jb);
        // ---- (1) panel step on rows jb .. h-1, columns jb .. jb+bw-1
        const int i0 = jb + t, i1 = jb + t + G;
        const bool l0 = i0 < h, l1 = i1 < h;
        const int nact = min(G, (h - jb + 31) & ~31);        // threads of the warps that own rows (second rows only exist when nact == G)
        if (t < nact) {
        double a0[kBlkNb], a1[kBlkNb];
#pragma unroll
        for (int k = 0; k < kBlkNb; ++k) {
            a0[k] = (l0 && k < bw) ? Ps[i0 + (jb + k) * h] : 0.0;
            a1[k] = (l1 && k < bw) ? Ps[i1 + (jb + k) * h] : 0.0;
        }
        if (t < bw) cb[t] = a0[0];                            // column jb of the pivot rows (row jb: the pivot), relative index
        blk_sync_active<G>(nact);
#pragma unroll
        for (int j = 0; j < kBlkNb; ++j) {
            if (j < bw) {
                const double* u = cb + (j & 1) * 8;
                double* un = cb + ((j + 1) & 1) * 8;
                const double d = u[0];
                bad |= !(fabs(d) > tiny && fabs(d) < 1.0e300);
                const double r = rr_rcp(d);
                const double m0 = a0[j] * r, m1 = a1[j] * r;
                if (j + 1 < kBlkNb) {
                    a0[j + 1 < kBlkNb ? j + 1 : 0] = fma(-m0, u[1], a0[j + 1 < kBlkNb ? j + 1 : 0]);
                    a1[j + 1 < kBlkNb ? j + 1 : 0] = fma(-m1, u[1], a1[j + 1 < kBlkNb ? j + 1 : 0]);
                    if (t > j && t < bw) un[t - j - 1] = a0[j + 1 < kBlkNb ? j + 1 : 0];   // pivot rows below the diagonal publish column j+1
                }
#pragma unroll
                for (int k = j + 2; k < kBlkNb; ++k) {
                    a0[k] = fma(-m0, u[k - j], a0[k]);
                    a1[k] = fma(-m1, u[k - j], a1[k]);
                }
                if (t == j) dblk[j] = d;
                if (t > j) { a0[j] = m0; if (growth > 0.0) bad |= l0 && !(fabs(m0) <= growth); }
                a1[j] = m1;
                if (growth > 0.0) bad |= l1 && !(fabs(m1) <= growth);
                blk_sync_active<G>(nact);
            }
        }
        // write the block back: D on the diagonal (row jb + j keeps d_j in a0[j]), L elsewhere
#pragma unroll
        for (int k = 0; k < kBlkNb; ++k) {
            if (k < bw) {
                if (l0 && t >= k) Ps[i0 + (jb + k) * h] = a0[k];
                if (l1) Ps[i1 + (jb + k) * h] = a1[k];
            }
        }
        }
        blk_sync<G>();
        // ---- (2) trailing update of rows / columns >= jb + 8 (lower triangle), 8 x 8 tiles on the tensor cores
        const int r00 = jb + kBlkNb;
        if (r00 < w) {
            const int nrt = (h - r00 + 7) >> 3, nct = (w - r00 + 7) >> 3;
            // tiles (ri, cj), cj <= min(ri, nct - 1), enumerated column by column: column cj holds nrt - cj tiles
            int q = gw, cj = 0, base = 0;
            const int ntile = nct * nrt - nct * (nct - 1) / 2;
            const double dk0 = fk < bw ? -dblk[fk] : 0.0, dk1 = fk + 4 < bw ? -dblk[fk + 4] : 0.0;
            __syncwarp();                                     // mma.sync.aligned below: the warp must be converged
            for (; q < ntile; q += NW) {
                while (q - base >= nrt - cj) { base += nrt - cj; ++cj; }
                const int ri = cj + (q - base);
                const int row = r00 + ri * 8 + fr, col = r00 + cj * 8 + 2 * fk;
                const int brow = r00 + cj * 8 + fr;                       // B operand: row (r00 + cj*8 + n) of L, n = fr
                const bool rin = row < h, bin = brow < h;
                double c0 = (rin && col < w) ? Ps[row + col * h] : 0.0;
                double c1 = (rin && col + 1 < w) ? Ps[row + (col + 1) * h] : 0.0;
                const double x0 = (rin && fk < bw) ? Ps[row + (jb + fk) * h] : 0.0;
                const double y0 = (bin && fk < bw) ? Ps[brow + (jb + fk) * h] * dk0 : 0.0;
                const double x1 = (rin && fk + 4 < bw) ? Ps[row + (jb + fk + 4) * h] : 0.0;
                const double y1 = (bin && fk + 4 < bw) ? Ps[brow + (jb + fk + 4) * h] * dk1 : 0.0;
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(x0), "d"(y0));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(x1), "d"(y1));
                if (rin && col < w && col <= row) Ps[row + col * h] = c0;
                if (rin && col + 1 < w && col + 1 <= row) Ps[row + (col + 1) * h] = c1;
            }
            blk_sync<G>();
        }
    }
    if (bad) atomicExch(fail, 1);
}

}  // namespace madipm
