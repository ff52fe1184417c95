// Numeric factorisation of the fronts NEAR THE ROOT of the supernodal elimination tree: one front per SM, one task per
// block and height, a grid barrier between two heights (same schedule, records and data flow as k_tree_top in
// kernels_tree.cuh, which this kernel replaces for OP = factor and factor + forward).
//
// Up there a height holds fewer fronts than the GPU has SMs, so a height costs the LATENCY of one front (~110 x 48 on the
// 1M-nnz QP), and the heights are sequential.  Measured on B200 with the blocked left-looking routine of kernels_tree.cuh:
// ~36 us per height, of which 20 in the dense LDL', 6.5 staging and pulling the children, 3.5 in the contribution block.
// What this kernel does differently:
//   * dense LDL' with the rows of the front in REGISTERS, right-looking, one barrier per column with look-ahead of the next
//     pivot column (dense_ldl_rr.cuh): half the cycles per column;
//   * the children's contribution blocks are read ONCE: the part that lands in the panel is added to it, the pass-through
//     part is accumulated in a shared-memory image of this front's own contribution block (k_tree adds it later with a
//     read-modify-write pass over global memory);
//   * the contribution block  C = pass-through - L21 D L21'  is computed on the FP64 tensor cores with TWO independent
//     accumulator chains per warp (a DMMA.8x8x4 has ~40 cycles of dependent latency) and stored once;
//   * the positions of the children's rows (par_pos) are staged in shared memory, so the scatter of a child's block costs
//     one dependent global round trip (the values), not two.
// Summation order differs from k_tree's (pass-through first, then the product), so results differ from that kernel at
// rounding level; they are bit-reproducible from run to run as before (no atomics, fixed order).
#pragma once
#include "dense_ldl_rr.cuh"
#include "kernels_tree.cuh"

namespace madipm {

constexpr int kFrontThreads = 256;
constexpr int kFrontPosCap = 4096;      // par_pos entries staged per chunk of children (ints)

// shared-memory layout (doubles), computed by the host (device.cu front_layout)
struct FrontLayout {
    int cap;        // panel
    int ctri;       // pass-through accumulator: packed lower triangle of order nbmax
    int maxh;       // su (forward step) / scratch
    int maxw;       // D, then y of the forward step
    int total;      // doubles of dynamic shared memory
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool FWD>
__device__ __forceinline__ void front_factor_task(const TreeView& T, const FrontLayout& Lo, const SnMeta& m, int inst, double* sm) {
    constexpr int G = kFrontThreads;
    constexpr int NW = G / 32;
    const FactorView& F = T.F;
    const int t = threadIdx.x;
    const int lane = t & 31, gw = t >> 5;
    const int c0 = m.c0, c1 = m.c1;
    const int w = m.w, h = m.h, nb = h - w;
    double* Lx = F.Lx + inst * F.lx_stride;
    double* P = Lx + m.xoff;
    double* Cx = T.Cx + inst * T.cx_stride;
    double* Ps = sm;
    double* Cacc = Ps + Lo.cap;
    double* ucol = Cacc + Lo.ctri;
    double* dk = ucol + Lo.maxh;
    double* cbuf = dk + Lo.maxw;                               // 2 * (64 + 2) doubles: column buffer of dense_ldl_rr
    ChildRec* cs = reinterpret_cast<ChildRec*>(cbuf + 2 * 66);
    int* pp = reinterpret_cast<int*>(cs + kChildChunk);        // kFrontPosCap ints
    const int np = h * w;
    const int nchunk = (np + 1) >> 1;
    const int ntri = nb * (nb + 1) / 2;

#ifdef MADIPM_TREE_TIMING
    long long ft[10];
#define FT(k) do { if (t == 0 && dbg_hit(T, m.sn)) ft[k] = clock64(); } while (0)
#else
#define FT(k)
#endif
    FT(0);
    __syncthreads();                                           // the previous task of this block is done with the scratch
    const double* Psrc = T.image ? T.image + inst * F.lx_stride + m.xoff : P;
    for (int c = t; c < nchunk; c += G) cp_async16(Ps + 2 * c, Psrc + 2 * c);
    const double* dpm = T.dperm ? T.dperm + inst * T.xw_stride + m.fc : nullptr;
    const double dg0 = (dpm && t < w) ? __ldcg(dpm + t) : 0.0;
    for (int e = t; e < ntri; e += G) Cacc[e] = 0.0;
    // 1. the children's contribution blocks, read once: entries of the first nbp columns land in the panel, the rest in Cacc
    bool waited = false;
    for (int base = 0; base < c1 - c0; base += kChildChunk) {
        const int n = stage_children<G>(T, cs, c0, c1, base, t);
        // stage the row positions of these children (as many as fit; the rest is read through the read-only path)
        {
            int o = 0;
            for (int ci = 0; ci < n; ++ci) {
                const int len = cs[ci].pull ? cs[ci].nb : 0;
                if (len > 0 && o + len <= kFrontPosCap)
                    for (int k = t; k < len; k += G) pp[o + k] = __ldg(T.par_pos + cs[ci].below + k);
                o += len;
            }
        }
        if (!waited) { cp_async_wait_all(); waited = true; }
        __syncthreads();
        int o = 0;
#pragma unroll 1
        for (int ci = 0; ci < n; ++ci) {
            const ChildRec r = cs[ci];
            if (!r.pull) continue;
            const bool staged = o + r.nb <= kFrontPosCap;
            const int* posg = T.par_pos + r.below;
            const int* poss = pp + o;
            o += r.nb;
            const double* Cc = Cx + r.coff;
            const int e1 = tri_colstart(r.nb, r.nb);
            constexpr int U = 8;
            for (int e0 = t; e0 < e1; e0 += U * G) {
                double v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = (e0 + u * G < e1) ? __ldcg(Cc + e0 + u * G) : 0.0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = e0 + u * G;
                    if (e < e1) {
                        const int b = tri_col(e, r.nb);
                        const int a = b + e - tri_colstart(b, r.nb);
                        const int pb = staged ? poss[b] : __ldg(posg + b);
                        const int pa = staged ? poss[a] : __ldg(posg + a);
                        if (pb < w) Ps[pa + pb * h] += v[u];
                        else Cacc[packed_col(pb - w, nb) + pa - w] += v[u];
                    }
                }
            }
            __syncthreads();                                   // the next child may hit the same entries
        }
    }
    if (!waited) { cp_async_wait_all(); __syncthreads(); }
    FT(1);
    if (dpm) {
        if (t < w) Ps[t * (h + 1)] += dg0;
        for (int k = t + G; k < w; k += G) Ps[k * (h + 1)] += __ldcg(dpm + k);
    }
    __syncthreads();
    // 2. dense LDL' (rows in registers); threads beyond the last row's warp skip it.  Pivot and growth checks ride along.
    {
        bool bad;
        switch ((w + 7) >> 3) {
            case 1: bad = dense_ldl_rr<8>(Ps, h, w, cbuf, t, 1, F.growth, F.tiny); break;
            case 2: bad = dense_ldl_rr<16>(Ps, h, w, cbuf, t, 1, F.growth, F.tiny); break;
            case 3: bad = dense_ldl_rr<24>(Ps, h, w, cbuf, t, 1, F.growth, F.tiny); break;
            case 4: bad = dense_ldl_rr<32>(Ps, h, w, cbuf, t, 1, F.growth, F.tiny); break;
            case 5: bad = dense_ldl_rr<40>(Ps, h, w, cbuf, t, 1, F.growth, F.tiny); break;
            case 6: bad = dense_ldl_rr<48>(Ps, h, w, cbuf, t, 1, F.growth, F.tiny); break;
            case 7: bad = dense_ldl_rr<56>(Ps, h, w, cbuf, t, 1, F.growth, F.tiny); break;
            default: bad = dense_ldl_rr<64>(Ps, h, w, cbuf, t, 1, F.growth, F.tiny); break;
        }
        if (bad) atomicExch(F.fail + inst, 1);
    }
    __syncthreads();
    FT(2);
    for (int c = t; c < nchunk; c += G) reinterpret_cast<double2*>(P)[c] = reinterpret_cast<const double2*>(Ps)[c];
    FT(3);
    if (nb > 0) {
        for (int k = t; k < w; k += G) dk[k] = Ps[k * (h + 1)];
        __syncthreads();
        // 3. C = Cacc - L21 D L21' on the FP64 tensor cores: the lower-triangular 8 x 8 blocks are dealt to the warps, two at a
        //    time per warp (independent accumulator chains); operands straight from the k-major panel, D applied on the fly
        double* Cs = Cx + m.coff;
        const int fr = lane >> 2, fk = lane & 3;
        const int nb8 = (nb + 7) >> 3;
        const int nblk = nb8 * (nb8 + 1) / 2;
        __syncwarp();                                          // mma.sync.aligned below: the warp must be converged
        for (int q0 = 2 * gw; q0 < nblk; q0 += 2 * NW) {
            int bi[2], bj[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {                      // block q of the row-major enumeration of the lower triangle
                const int q = min(q0 + u, nblk - 1);
                int i = static_cast<int>((sqrtf(8.0f * q + 1.0f) - 1.0f) * 0.5f);
                while ((i + 1) * (i + 2) / 2 <= q) ++i;
                while (i * (i + 1) / 2 > q) --i;
                bi[u] = i;
                bj[u] = q - i * (i + 1) / 2;
            }
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            const int ra0 = w + bi[0] * 8 + fr, rb0 = w + bj[0] * 8 + fr, ra1 = w + bi[1] * 8 + fr, rb1 = w + bj[1] * 8 + fr;
            for (int k0 = 0; k0 < w; k0 += 4) {
                const int kk = k0 + fk;
                const bool kin = kk < w;
                const double dkk = kin ? dk[kk] : 0.0;
                const double a0 = (kin && ra0 < h) ? Ps[ra0 + kk * h] : 0.0;
                const double b0 = (kin && rb0 < h) ? Ps[rb0 + kk * h] * dkk : 0.0;
                const double a1 = (kin && ra1 < h) ? Ps[ra1 + kk * h] : 0.0;
                const double b1 = (kin && rb1 < h) ? Ps[rb1 + kk * h] * dkk : 0.0;
                dmma884(acc[0][0], acc[0][1], a0, b0);
                dmma884(acc[1][0], acc[1][1], a1, b1);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (q0 + u >= nblk) break;
                const int arow = bi[u] * 8 + fr, col0 = bj[u] * 8 + 2 * fk;
                if (arow < nb) {
                    if (arow >= col0) { const int o = packed_col(col0, nb) + arow; Cs[o] = Cacc[o] - acc[u][0]; }
                    if (arow >= col0 + 1 && col0 + 1 < nb) { const int o = packed_col(col0 + 1, nb) + arow; Cs[o] = Cacc[o] - acc[u][1]; }
                }
            }
        }
    }
    FT(4);
    if (FWD) {
        __syncthreads();                                       // everyone is done with D (dk) and the staged child records
        tree_forward_core<G>(T, m, inst, Ps, ucol, dk, cs, t);
    }
    FT(5);
#ifdef MADIPM_TREE_TIMING
    if (t == 0 && dbg_hit(T, m.sn))
        printf("front sn %d w %d h %d children %d | stage+children %lld ldl %lld check+store %lld schur %lld fwd %lld cycles\n", m.sn, w, h,
               c1 - c0, ft[1] - ft[0], ft[2] - ft[1], ft[3] - ft[2], ft[4] - ft[3], ft[5] - ft[4]);
#endif
}

// OP: 0 factor, 3 factor + forward step.  Heights ph_begin .. ph_end - 1, each with at most gridDim.x * k tasks.
template <int OP>
__global__ void __launch_bounds__(kFrontThreads, 1)
k_front(TreeView T, FrontLayout Lo, int ph_begin, int ph_end) {
    extern __shared__ double2 tree_smem2[];
    double* smem = reinterpret_cast<double*>(tree_smem2);
    unsigned target = 0;
    for (int ph = ph_begin; ph < ph_end; ++ph) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(&T.dbg->progress) = (2u << 24) | (OP << 16) | ph;
        const int first = __ldg(T.phase_cls + ph * 5);
        const int n = __ldg(T.phase_cls + ph * 5 + 4) - first;
        const int64_t total = static_cast<int64_t>(n) * T.B;
        for (int64_t q = blockIdx.x; q < total; q += gridDim.x) {
            const int inst = static_cast<int>(q / n);
            front_factor_task<OP == 3>(T, Lo, ldg_rec(T.meta + first + static_cast<int>(q - static_cast<int64_t>(inst) * n)), inst, smem);
        }
        if (ph + 1 != ph_end && !grid_barrier(T, target, 2, OP, ph)) return;
    }
}

}  // namespace madipm
