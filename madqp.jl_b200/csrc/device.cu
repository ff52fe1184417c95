// Device half of the C ABI: device state, uploads and every numeric entry point.
// One handle = one CUDA stream; all work is stream-ordered; entry points that return
// scalars synchronise, the others do not.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "handle.hpp"
#include "kernels_factor.cuh"
#include "kernels_tree.cuh"
#include "kernels_front.cuh"
#include "kernels_big.cuh"
#include "kernels_vec.cuh"

namespace madipm {

struct DeviceState {
    int device = 0;
    cudaStream_t stream = nullptr;
    int B = 1;
    int n = 0, m = 0, nlb = 0, nub = 0, N = 0, kind = 0;
    bool scaled = false;               // K2.5 (MADIPM_KKT_K2_SCALED): kind == 0 pattern, values rescaled every iteration
    int *kkt_row = nullptr, *kkt_col = nullptr;   // scaled only: coordinates of every nonzero of aug_com
    double* nz_base = nullptr;                    // scaled only: the unscaled constant part of aug_com.nzval
    int64_t L = 0;
    int64_t nnzA = 0, nnzQ = 0, nnzh = 0, nnzj = 0, nnzJfull = 0, kkt_nnz = 0, lx_stride = 0;
    int nsuper = 0, n_levels = 0;
    int tri_cap = 0;
    size_t smem_tri = 0;
    // small supernodes: persistent height-scheduled kernels (kernels_tree.cuh)
    // heights [ph0, split) in one cooperative launch of k_tree, heights [split, ph1) (at most one task per SM) in one
    // launch of k_tree_top; a height with big supernodes ends a segment
    struct TreeSeg { int ph0, split, ph1, grid[4], grid_top; bool df_sweeps; };   // grid[OP], OP 3 = factor + forward (kernels_tree.cuh)
    std::vector<TreeSeg> segs;
    int tree_maxw = 0;
    size_t smem_tree = 0;
    int tree_occ[4] = {0, 0, 0, 0};
    int* tickets[3] = {nullptr, nullptr, nullptr};
    // dataflow scheduling of the k_tree launches (kernels_tree.cuh, DfView): dependency counters instead of height barriers
    bool dataflow = false;             // on by default (device_create), MADIPM_TREE_DATAFLOW=0 / madipm_set_tree_dataflow switch it off; only without multi-block fronts and per-height launches
    bool dataflow_ok = false;          // the schedule allows it
    int* task_parent = nullptr;        // [ntasks]
    int* dep = nullptr;                // [batch][ntasks]
    int ntasks = 0;
    std::vector<int> phase_ptr;        // first task of every height (+ end)
    int tree_occ_df[4] = {0, 0, 0, 0};
    bool use_front = false;            // the heights near the root are factorised by k_front (kernels_front.cuh) instead of k_tree_top
    FrontLayout front_lo = {};
    size_t smem_front = 0;
    int* phase_cls = nullptr;
    int num_sms = 0;
    unsigned* abort_flag = nullptr;    // device: a grid barrier gave up (kernels_tree.cuh grid_barrier)
    TreeDebug* tree_dbg = nullptr;     // mapped pinned host memory: what that barrier saw
    unsigned aborts_seen = 0;          // value of tree_dbg->aborted already reported
    double sync_timeout_s = 120.0;     // host-side limit for one stream synchronisation (MADIPM_SYNC_TIMEOUT_S)
    long long spin_budget_ns = 2000000000ll;   // device-side limit for one block at one grid barrier (MADIPM_BARRIER_TIMEOUT_MS)
    int* par_pos = nullptr;
    SnMeta* sn_meta = nullptr;
    ChildRec* child_rec = nullptr;
    double *Cx = nullptr, *cvec = nullptr;
    int64_t cx_stride = 0, cv_stride = 0;
    std::vector<int> big_ptr;          // big (multi-block) supernodes, per height
    std::vector<int> big_rowblocks, big_rowblocks512, big_tiles, big_panelblocks;   // per height: grid sizes (max over its big supernodes)
    bool big_use_panel = false;
    size_t smem_big_panel = 0;
    int* big_sn = nullptr;
    int *upd_lim = nullptr, *upd_nsrc = nullptr, *upd_src = nullptr;   // per entry of big_sn: see the grouping in device_create
    double *big_tbuf = nullptr, *big_zbuf = nullptr;
    int* big_done = nullptr;           // per (instance, big supernode of a height): blocks finished (self-resetting)
    int big_zstride = 0;
    int big_max_per_level = 0;
    size_t smem_big_diag = 0, smem_big_trsm = 0, smem_big_update = 0;

    std::vector<void*> allocs;

    int *ind_lb = nullptr, *ind_ub = nullptr, *lb_of = nullptr, *ub_of = nullptr;
    int *A_rp = nullptr, *A_ci = nullptr, *AT_rp = nullptr, *AT_ci = nullptr, *Q_rp = nullptr, *Q_ci = nullptr;
    int *A_order = nullptr, *AT_src = nullptr, *Q_src = nullptr;
    double *A_val = nullptr, *AT_val = nullptr, *Q_val = nullptr, *hessV = nullptr, *jacV = nullptr;
    int *aug_map = nullptr, *diag_slot = nullptr, *jac_row = nullptr;
    double* d_objscale = nullptr;
    int nx = 0, n_slack = 0;
    double* const_diag = nullptr;
    int64_t* prod_ptr = nullptr;
    int *prod_a = nullptr, *prod_b = nullptr, *prod_k = nullptr;

    int *sn_first = nullptr, *sn_rows = nullptr, *sn_of = nullptr;
    int64_t *sn_rowptr = nullptr, *sn_xoff = nullptr, *asm_map = nullptr, *diag_pos = nullptr;
    int64_t *row_colbase = nullptr, *row_relbase = nullptr;
    int* rel = nullptr;
    double *Lx = nullptr, *xw = nullptr;
    // K2 without big supernodes: the assembled panels are kept as an image P0 (constant part written once per
    // set_values, the N diagonal slots rewritten by build_kkt); the factor tasks read P0 and write Lx, so a
    // factorisation needs neither the memset of Lx nor k_assemble_panels.
    double* P0 = nullptr;
    int64_t* diag_lx = nullptr;
    bool image_valid = false;
    // With a valid image, build_kkt writes only `dperm` = the per-iteration diagonals (pr_diag | du_diag) in PIVOT order
    // (one coalesced pass); the factor task of a supernode adds its w entries to the diagonal of the staged panel.
    // The diagonal slots of KKT_NZVAL are then brought up to date only when somebody looks at that vector.
    double* dperm = nullptr;
    bool nzval_diag_stale = false;
    int *perm = nullptr, *iperm = nullptr;
    int* fail = nullptr;
    std::vector<int> h_fail;
    // madipm_factorize(h, NULL) does not synchronise: the breakdown flags (and, with light profiling, the event times)
    // are collected by the next call that synchronises anyway (any fetch, madipm_is_factorized, madipm_synchronize)
    int* h_fail_pinned = nullptr;
    bool fail_pending = false, prof_pending = false, prof_tree_pending = false;
    double *d_norms = nullptr, *d_obj = nullptr;   // outputs of the fused phases that share one synchronisation
    double *d_iter = nullptr, *h_iter = nullptr;   // madipm_phase_iteration: [aff4 4B | norms2 3B | ratio8 8B | res4 4B] device, 28B pinned
    int* d_ok = nullptr;

    double* vec[MADIPM_VEC_COUNT] = {};
    int64_t veclen[MADIPM_VEC_COUNT] = {};
    double *buffer_n = nullptr, *buffer_m = nullptr;
    double* snap[4] = {nullptr, nullptr, nullptr, nullptr};   // device copies of X, XL, XU, Y

    double* red_scratch = nullptr;
    unsigned int* red_counter = nullptr;
    double *d_out = nullptr, *d_out2 = nullptr, *d_steps = nullptr, *d_mu = nullptr, *d_tau = nullptr, *d_ones = nullptr;
    int* d_active = nullptr;
    unsigned long long* d_inertia = nullptr;
    double* h_out = nullptr;   // pinned

    const Symbolic* sym = nullptr;
    bool values_set = false;
    double growth_limit = 1.0e13, tiny_pivot = 1.0e-13;   // see madipm_set_pivot_checks
    bool inject_barrier_fault = false; // test hook: see madipm_debug_inject_barrier_fault
    bool wedged = false;               // a synchronisation timed out: the stream still holds a kernel that never returns
    int profiling = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr;   // ev4: before the top launch of the first segment
    double t_ms[MADIPM_T_COUNT] = {};
    int64_t t_calls[MADIPM_T_COUNT] = {};
};

namespace {

#define CUDA_OK(h, call)                                                                   \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            (h)->error = std::string(#call) + ": " + cudaGetErrorString(e_);               \
            return MADIPM_CUDA_ERROR;                                                      \
        }                                                                                  \
    } while (0)

#define LAUNCH(h, kernel, grid, block, smem, ...)                                          \
    do {                                                                                   \
        kernel<<<grid, block, smem, (h)->dev->stream>>>(__VA_ARGS__);                      \
        ++(h)->launches;                                                                   \
        (h)->last_launch = #kernel;                                                        \
    } while (0)

template <class T>
int32_t dev_alloc(madipm_handle* h, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    void* q = nullptr;
    CUDA_OK(h, cudaMalloc(&q, count * sizeof(T)));
    h->dev->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return MADIPM_OK;
}

template <class T, class S>
int32_t upload_as(madipm_handle* h, T** p, const std::vector<S>& v) {
    std::vector<T> tmp(v.size());
    for (size_t i = 0; i < v.size(); ++i) tmp[i] = static_cast<T>(v[i]);
    int32_t rc = dev_alloc(h, p, tmp.size());
    if (rc) return rc;
    if (!tmp.empty()) CUDA_OK(h, cudaMemcpy(*p, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
    return MADIPM_OK;
}

inline int vec_blocks(int64_t n) {
    int64_t b = (n + kVecThreads - 1) / kVecThreads;
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(b, kMaxRedBlocks)));
}

VecView make_view(const DeviceState& d) {
    VecView V;
    V.n = d.n; V.m = d.m; V.nlb = d.nlb; V.nub = d.nub;
    V.ind_lb = d.ind_lb; V.ind_ub = d.ind_ub; V.lb_of = d.lb_of; V.ub_of = d.ub_of;
    V.x = d.vec[MADIPM_VEC_X]; V.xl = d.vec[MADIPM_VEC_XL]; V.xu = d.vec[MADIPM_VEC_XU];
    V.zl = d.vec[MADIPM_VEC_ZL]; V.zu = d.vec[MADIPM_VEC_ZU]; V.f = d.vec[MADIPM_VEC_F];
    V.jacl = d.vec[MADIPM_VEC_JACL]; V.cost = d.vec[MADIPM_VEC_COST];
    V.y = d.vec[MADIPM_VEC_Y]; V.c = d.vec[MADIPM_VEC_C]; V.rhs = d.vec[MADIPM_VEC_RHS];
    V.reg = d.vec[MADIPM_VEC_REG]; V.pr_diag = d.vec[MADIPM_VEC_PR_DIAG]; V.du_diag = d.vec[MADIPM_VEC_DU_DIAG];
    V.l_diag = d.vec[MADIPM_VEC_L_DIAG]; V.l_lower = d.vec[MADIPM_VEC_L_LOWER];
    V.u_diag = d.vec[MADIPM_VEC_U_DIAG]; V.u_lower = d.vec[MADIPM_VEC_U_LOWER];
    V.corr_lb = d.vec[MADIPM_VEC_CORR_LB]; V.corr_ub = d.vec[MADIPM_VEC_CORR_UB];
    V.scaling = d.vec[MADIPM_VEC_SCALING]; V.dsg = d.scaled ? -1.0 : 1.0;
    V.sn = d.n; V.sm = d.m; V.slb = d.nlb; V.sub = d.nub; V.sL = d.L;
    return V;
}

FactorView make_factor_view(const DeviceState& d, int = 0) {
    FactorView F;
    F.sn_first = d.sn_first; F.sn_rowptr = d.sn_rowptr; F.sn_rows = d.sn_rows; F.sn_xoff = d.sn_xoff;
    F.sn_of = d.sn_of; F.row_colbase = d.row_colbase; F.row_relbase = d.row_relbase; F.rel = d.rel;
    F.Lx = d.Lx; F.lx_stride = d.lx_stride; F.fail = d.fail; F.growth = d.growth_limit; F.tiny = d.tiny_pivot;
    return F;
}

TreeView make_tree_view(const DeviceState& d) {
    TreeView T;
    T.F = make_factor_view(d, 0);
    T.meta = d.sn_meta; T.crec = d.child_rec; T.par_pos = d.par_pos;
    T.Cx = d.Cx; T.cx_stride = d.cx_stride; T.cvec = d.cvec; T.cv_stride = d.cv_stride;
    T.image = nullptr; T.dperm = nullptr;
    T.xw = d.xw; T.xw_stride = d.N; T.bar = nullptr; T.B = d.B; T.tickets = nullptr; T.phase_cls = d.phase_cls;
    T.abort_flag = d.abort_flag; T.dbg = d.tree_dbg; T.spin_budget_ns = d.spin_budget_ns;
    for (int c = 0; c < 4; ++c) { T.slice[c] = d.sym->tree_slice[c]; T.cap[c] = d.sym->tree_cap[c]; T.maxh[c] = d.sym->tree_maxh[c]; }
    T.maxw = d.tree_maxw;
    {   // MADIPM_DBG_SN=a,b,...: supernodes whose stage clocks a MADIPM_TREE_TIMING build prints
        for (int k = 0; k < 8; ++k) T.dbg_sn[k] = -1;
        const char* e = std::getenv("MADIPM_DBG_SN");
        for (int k = 0; e && *e && k < 8; ++k) {
            T.dbg_sn[k] = std::atoi(e);
            while (*e && *e != ',') ++e;
            if (*e == ',') ++e;
        }
    }
    return T;
}

struct Timer {
    madipm_handle* h;
    int cls;
    Timer(madipm_handle* hh, int c) : h(hh), cls(c) {
        if (h->dev->profiling == 1) cudaEventRecord(h->dev->ev0, h->dev->stream);
    }
    ~Timer() {
        DeviceState* d = h->dev;
        if (d->profiling == 1) {
            cudaEventRecord(d->ev1, d->stream);
            cudaEventSynchronize(d->ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, d->ev0, d->ev1);
            d->t_ms[cls] += ms;
            d->t_calls[cls] += 1;
        }
    }
};

// Every synchronisation of the handle's stream goes through here.  Two guards (a grid that never returned
// pinned a GPU for 30 minutes without a word):
//   * the wait is bounded (MADIPM_SYNC_TIMEOUT_S, default 120 s, 0 = plain cudaStreamSynchronize): on expiry the call
//     returns MADIPM_INTERNAL with whatever the kernels left in the host-visible debug record, and the handle is marked
//     wedged (its destructor then leaks instead of blocking in cudaFree);
//   * a grid barrier that gave up on the device (bounded spin, kernels_tree.cuh) surfaces here as MADIPM_INTERNAL.
std::string describe_tree_debug(const DeviceState* d) {
    const TreeDebug g = d->tree_dbg ? *d->tree_dbg : TreeDebug{};
    if (!g.aborted) return "no grid barrier reported a timeout";
    static const char* kn[] = {"k_tree", "k_tree_top", "k_front", "k_tree_df"};
    static const char* on[] = {"factor", "forward", "backward"};
    char buf[320];
    if ((g.kernel & 3) == 3)
        std::snprintf(buf, sizeof(buf), "dependency wait timed out: k_tree_df<%s> in height %d, block %u of %u saw %u of %u signals, for %.0f ms",
                      on[g.op == 3 ? 0 : g.op % 3], g.phase, g.block, g.grid, g.seen, g.target, g.waited_ns / 1e6);
    else
    std::snprintf(buf, sizeof(buf), "grid barrier timed out: %s<%s> after height %d, block %u of %u saw %u arrivals, waited for %u, for %.0f ms",
                  kn[g.kernel & 3], on[g.op == 3 ? 0 : g.op % 3], g.phase, g.block, g.grid, g.seen, g.target, g.waited_ns / 1e6);
    return buf;
}

int32_t sync_stream(madipm_handle* h) {
    DeviceState* d = h->dev;
    if (d->wedged) { h->error = "handle is wedged by an earlier synchronisation timeout"; return MADIPM_INTERNAL; }
    cudaError_t e;
    if (d->sync_timeout_s <= 0.0) {
        e = cudaStreamSynchronize(d->stream);
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        unsigned polls = 0;
        while ((e = cudaStreamQuery(d->stream)) == cudaErrorNotReady) {
            if ((++polls & 0xfffu) == 0 &&
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > d->sync_timeout_s) {
                char buf[128];
                std::snprintf(buf, sizeof(buf), "stream still busy after %.0f s (MADIPM_SYNC_TIMEOUT_S); ", d->sync_timeout_s);
                char prog[160];
                const unsigned pg = d->tree_dbg ? d->tree_dbg->progress : 0u;
                std::snprintf(prog, sizeof(prog), "; last kernel launched: %s (launch %lld); last tree progress mark: kernel %u op %u height %u",
                              h->last_launch, static_cast<long long>(h->launches), pg >> 24, (pg >> 16) & 0xffu, pg & 0xffffu);
                h->error = std::string(buf) + describe_tree_debug(d) + prog;
                d->wedged = true;
                return MADIPM_INTERNAL;
            }
        }
    }
    if (e != cudaSuccess) { h->error = std::string("stream synchronisation: ") + cudaGetErrorString(e); return MADIPM_CUDA_ERROR; }
    if (d->tree_dbg && d->tree_dbg->aborted != d->aborts_seen) {
        d->aborts_seen = d->tree_dbg->aborted;
        h->error = describe_tree_debug(d);
        cudaMemsetAsync(d->abort_flag, 0, sizeof(unsigned), d->stream);      // the next launch starts clean
        cudaStreamSynchronize(d->stream);
        return MADIPM_INTERNAL;
    }
    return MADIPM_OK;
}

// what an asynchronous factorisation left behind: enqueue before a synchronisation, collect after it
int32_t pending_enqueue(madipm_handle* h) {
    DeviceState* d = h->dev;
    if (d->fail_pending)
        CUDA_OK(h, cudaMemcpyAsync(d->h_fail_pinned, d->fail, sizeof(int) * d->B, cudaMemcpyDeviceToHost, d->stream));
    return MADIPM_OK;
}

void pending_collect(madipm_handle* h) {
    DeviceState* d = h->dev;
    if (d->fail_pending) {
        for (int i = 0; i < d->B; ++i) d->h_fail[i] = d->h_fail_pinned[i];
        d->fail_pending = false;
    }
    if (d->prof_pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, d->ev0, d->ev1) == cudaSuccess) { d->t_ms[MADIPM_T_FACTOR] += ms; d->t_calls[MADIPM_T_FACTOR] += 1; }
        if (d->prof_tree_pending && cudaEventElapsedTime(&ms, d->ev2, d->ev3) == cudaSuccess) {
            d->t_ms[MADIPM_T_FACTOR_TREE] += ms; d->t_calls[MADIPM_T_FACTOR_TREE] += 1;
            if (cudaEventElapsedTime(&ms, d->ev4, d->ev3) == cudaSuccess) { d->t_ms[MADIPM_T_FACTOR_TOP] += ms; d->t_calls[MADIPM_T_FACTOR_TOP] += 1; }
        }
        d->prof_pending = d->prof_tree_pending = false;
    }
}

int32_t fetch(madipm_handle* h, const double* dsrc, double* hdst, size_t count) {
    DeviceState* d = h->dev;
    int32_t rc;
    if ((rc = pending_enqueue(h))) return rc;
    CUDA_OK(h, cudaMemcpyAsync(d->h_out, dsrc, count * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
    { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }
    CUDA_OK(h, cudaGetLastError());
    pending_collect(h);
    std::memcpy(hdst, d->h_out, count * sizeof(double));
    return MADIPM_OK;
}

int32_t spmv_launch(madipm_handle* h, int rows, int64_t nnz, const int* rp, const int* ci, const double* val,
                    int64_t val_stride, const double* x, int64_t xs, double* y, int64_t ys, double alpha, double beta) {
    DeviceState* d = h->dev;
    if (rows <= 0) return MADIPM_OK;
    const double avg = static_cast<double>(nnz) / rows;
    dim3 grid(1, d->B);
    auto blocks = [&](int lanes) {
        int64_t rpb = static_cast<int64_t>(kVecThreads / lanes) * kSpmvRows;
        return static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>((rows + rpb - 1) / rpb, 148 * 16)));
    };
    if (avg <= 2.5) { grid.x = blocks(2); LAUNCH(h, k_spmv<2>, grid, kVecThreads, 0, rows, rp, ci, val, val_stride, x, xs, y, ys, alpha, beta); }
    else if (avg <= 6.0) { grid.x = blocks(4); LAUNCH(h, k_spmv<4>, grid, kVecThreads, 0, rows, rp, ci, val, val_stride, x, xs, y, ys, alpha, beta); }
    else if (avg <= 12.0) { grid.x = blocks(8); LAUNCH(h, k_spmv<8>, grid, kVecThreads, 0, rows, rp, ci, val, val_stride, x, xs, y, ys, alpha, beta); }
    else if (avg <= 24.0) { grid.x = blocks(16); LAUNCH(h, k_spmv<16>, grid, kVecThreads, 0, rows, rp, ci, val, val_stride, x, xs, y, ys, alpha, beta); }
    else { grid.x = blocks(32); LAUNCH(h, k_spmv<32>, grid, kVecThreads, 0, rows, rp, ci, val, val_stride, x, xs, y, ys, alpha, beta); }
    return MADIPM_OK;
}

__global__ void k_scatter_add(int64_t n, const int* __restrict__ map, const double* __restrict__ src, int64_t sstride,
                              double* __restrict__ dst, int64_t dstride) {
    src += blockIdx.y * sstride;
    dst += blockIdx.y * dstride;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x)
        atomicAdd(&dst[map[k]], src[k]);
}

// K2a: only the diagonal slots of aug_com.nzval change between iterations
// (transfer! semantics, /root/reference/scripts/cuda_wrapper.jl:17-28, restricted to the entries that moved)
__global__ void k_build_kkt_k2(int n, int m, const int* __restrict__ diag_slot, const double* __restrict__ const_diag,
                               const double* __restrict__ pr_diag, const double* __restrict__ du_diag,
                               double* __restrict__ nzval, int64_t nz_stride) {
    const int inst = blockIdx.y;
    const_diag += static_cast<int64_t>(inst) * (n + m);
    pr_diag += static_cast<int64_t>(inst) * n;
    du_diag += static_cast<int64_t>(inst) * m;
    nzval += inst * nz_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n + m; i += gridDim.x * blockDim.x)
        nzval[diag_slot[i]] = ((i < n) ? pr_diag[i] : du_diag[i - n]) + const_diag[i];
}

// K2.5 build_kkt!: _build_scale_augmented_system_coo! (/root/reference/scripts/cuda_wrapper.jl:90-116) followed by
// transfer! (:17-28), restated per nonzero of aug_com: Hessian block base * s[row] * s[col], Jacobian block
// base * s[col], diagonal slots + pr_diag / du_diag (unscaled).  `base` holds the constant part with duplicates merged.
__global__ void k_build_kkt_k25(int64_t nnz, int n, int m, const int* __restrict__ row, const int* __restrict__ col,
                                const double* __restrict__ base, const double* __restrict__ scaling,
                                const double* __restrict__ pr_diag, const double* __restrict__ du_diag,
                                double* __restrict__ nzval) {
    const int inst = blockIdx.y;
    base += inst * nnz; nzval += inst * nnz;
    scaling += static_cast<int64_t>(inst) * n;
    pr_diag += static_cast<int64_t>(inst) * n;
    du_diag += static_cast<int64_t>(inst) * m;
    for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < nnz; k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int i = row[k], j = col[k];
        double v = base[k];
        if (i < n) v = __dmul_rn(__dmul_rn(v, scaling[i]), scaling[j]);
        else if (j < n) v = __dmul_rn(v, scaling[j]);
        if (i == j) v = __dadd_rn(v, i < n ? pr_diag[i] : du_diag[i - n]);
        nzval[k] = v;
    }
}

// K2a with a panel image: dperm[k] = (pr_diag | du_diag)[perm[k]], k in pivot order (coalesced writes, gathers from
// two vectors that stay in L2)
__global__ void k_build_kkt_k2_perm(int n, int m, const int* __restrict__ perm, const double* __restrict__ pr_diag,
                                    const double* __restrict__ du_diag, double* __restrict__ dperm) {
    const int inst = blockIdx.y;
    pr_diag += static_cast<int64_t>(inst) * n;
    du_diag += static_cast<int64_t>(inst) * m;
    dperm += static_cast<int64_t>(inst) * (n + m);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n + m; k += gridDim.x * blockDim.x) {
        const int i = __ldg(perm + k);
        dperm[k] = (i < n) ? pr_diag[i] : du_diag[i - n];
    }
}

__global__ void k_extract_diag(int N, const int* __restrict__ diag_slot, const double* __restrict__ nzval,
                               int64_t nz_stride, double* __restrict__ const_diag) {
    nzval += blockIdx.y * nz_stride;
    const_diag += static_cast<int64_t>(blockIdx.y) * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
        const_diag[i] = nzval[diag_slot[i]];
}

__global__ void k_perm_gather(int N, const int* __restrict__ perm, const double* __restrict__ src, int64_t sstride,
                              double* __restrict__ dst, int64_t dstride) {
    src += blockIdx.y * sstride;
    dst += blockIdx.y * dstride;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) dst[k] = src[perm[k]];
}

__global__ void k_perm_scatter(int N, const int* __restrict__ perm, const double* __restrict__ src, int64_t sstride,
                               double* __restrict__ dst, int64_t dstride) {
    src += blockIdx.y * sstride;
    dst += blockIdx.y * dstride;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) dst[perm[k]] = src[k];
}

int32_t check(madipm_handle* h, bool need_values = false) {
    if (!h) return MADIPM_BAD_ARGUMENT;
    if (!h->analyzed) { h->error = "numeric call before analyze"; return MADIPM_BAD_ARGUMENT; }
    if (!h->dev) {
        int32_t rc = device_create(h);
        if (rc) { device_destroy(h); return rc; }
    }
    if (need_values && !h->dev->values_set) { h->error = "numeric call before set_values"; return MADIPM_BAD_ARGUMENT; }
    cudaSetDevice(h->dev->device);
    return MADIPM_OK;
}

template <int OP>
int32_t launch_tree(madipm_handle* h, const DeviceState::TreeSeg& sg) {
    DeviceState* d = h->dev;
    TreeView T = make_tree_view(*d);
    int* const tickets = d->tickets[OP == 3 ? 0 : OP];     // (the fused factor + forward launch never overlaps a plain factor launch)
    T.tickets = tickets;
    if ((OP == 0 || OP == 3) && d->image_valid) { T.image = d->P0; T.dperm = d->dperm; }
    const bool reverse = OP == 2;
    // one memset clears everything the two launches of this segment synchronise on: the tickets of its heights and the
    // barrier counters (slot kSyncBar of the first height each launch walks)
    if (sg.ph1 > sg.ph0)
        CUDA_OK(h, cudaMemsetAsync(tickets + kSyncStride * sg.ph0, 0, sizeof(int) * kSyncStride * (sg.ph1 - sg.ph0), d->stream));
    const bool df_fault = d->inject_barrier_fault && d->dataflow && d->dataflow_ok && (OP == 0 || OP == 3 || sg.df_sweeps);   // ... or no dependency counter can
    if (d->inject_barrier_fault) {     // counters start at 0x80000000: no barrier of this segment can complete
        d->inject_barrier_fault = false;
        for (int ph : {sg.ph0, sg.split})
            if (ph < sg.ph1)
                CUDA_OK(h, cudaMemsetAsync(reinterpret_cast<char*>(tickets + kSyncStride * ph + kSyncBar) + 3, 0x80, 1, d->stream));
    }
    auto bottom = [&]() -> int32_t {
        if (sg.split <= sg.ph0 || sg.grid[OP] <= 0) return MADIPM_OK;
        int b = reverse ? sg.split - 1 : sg.ph0, e = reverse ? sg.ph0 - 1 : sg.split, step = reverse ? -1 : 1;
        T.bar = reinterpret_cast<unsigned*>(tickets + kSyncStride * sg.ph0 + kSyncBar);
        if (d->dataflow && d->dataflow_ok && (OP == 0 || OP == 3 || sg.df_sweeps)) {   // dependency counters instead of height barriers (kernels_tree.cuh, DfView)
            DfView D{d->dep, d->task_parent, d->ntasks, d->phase_ptr[sg.split]};
            CUDA_OK(h, cudaMemsetAsync(d->dep, df_fault ? 0x80 : 0, sizeof(int) * static_cast<size_t>(d->B) * d->ntasks, d->stream));
            const int grid = std::min(sg.grid[OP], d->tree_occ_df[OP] * d->num_sms);
            void* dargs[] = {&T, &D, &b, &e, &step};
            CUDA_OK(h, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_tree_df<OP>), dim3(grid), dim3(kTreeThreads), dargs,
                                                   d->smem_tree, d->stream));
            ++h->launches;
            h->last_launch = "k_tree_df (cooperative)";
            return MADIPM_OK;
        }
        void* args[] = {&T, &b, &e, &step};
        CUDA_OK(h, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_tree<OP>), dim3(sg.grid[OP]), dim3(kTreeThreads), args,
                                               d->smem_tree, d->stream));
        ++h->launches;
        h->last_launch = "k_tree (cooperative)";
        return MADIPM_OK;
    };
    auto top = [&]() -> int32_t {
        if (sg.ph1 <= sg.split || sg.grid_top <= 0) return MADIPM_OK;
        int b = reverse ? sg.ph1 - 1 : sg.split, e = reverse ? sg.split - 1 : sg.ph1, step = reverse ? -1 : 1;
        T.bar = reinterpret_cast<unsigned*>(tickets + kSyncStride * sg.split + kSyncBar);
        if ((OP == 0 || OP == 3) && d->use_front) {
            FrontLayout lo = d->front_lo;
            void* fargs[] = {&T, &lo, &b, &e};
            CUDA_OK(h, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_front<OP == 3 ? 3 : 0>), dim3(sg.grid_top), dim3(kFrontThreads),
                                                   fargs, d->smem_front, d->stream));
            ++h->launches;
            h->last_launch = "k_front (cooperative)";
            return MADIPM_OK;
        }
        void* args[] = {&T, &b, &e, &step};
        CUDA_OK(h, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_tree_top<OP>), dim3(sg.grid_top), dim3(kTopThreads), args,
                                               d->smem_tree, d->stream));
        ++h->launches;
        h->last_launch = "k_tree_top (cooperative)";
        return MADIPM_OK;
    };
    int32_t rc;
    if (reverse) { if ((rc = top())) return rc; return bottom(); }
    if ((rc = bottom())) return rc;
    if ((OP == 0 || OP == 3) && d->profiling == 2 && &sg == &d->segs.front()) cudaEventRecord(d->ev4, d->stream);
    return top();
}

// do_forward = false: the forward sweep of this right-hand side was done by the fused factor + forward launch
int32_t tri_solve(madipm_handle* h, bool do_forward = true) {
    DeviceState* d = h->dev;
    FactorView F = make_factor_view(*d, 0);
    const int64_t N = d->N;
    int32_t rc;
    for (const DeviceState::TreeSeg& sg : d->segs) {
        if (!do_forward) break;
        if ((rc = launch_tree<1>(h, sg))) return rc;
        const int l = sg.ph1 - 1;
        const int nbig = d->big_ptr[l + 1] - d->big_ptr[l];
        if (nbig > 0) {
            const int* list = d->big_sn + d->big_ptr[l];
            LAUNCH(h, k_big_forward, dim3(std::max(1, d->big_rowblocks[l] / 2 + d->big_rowblocks[l] % 2), nbig, d->B), 128, d->smem_tri, F, list,
                   d->xw, N, d->tri_cap, d->big_done);
        }
    }
    for (auto it = d->segs.rbegin(); it != d->segs.rend(); ++it) {
        const int l = it->ph1 - 1;
        const int nbig = d->big_ptr[l + 1] - d->big_ptr[l];
        if (nbig > 0) {
            const int* list = d->big_sn + d->big_ptr[l];
            LAUNCH(h, k_big_backward, dim3(std::max(1, d->big_rowblocks512[l]), nbig, d->B), 256, d->smem_tri, F, list, d->big_tbuf, 256,
                   d->xw, N, d->tri_cap, d->big_done);
        }
        if ((rc = launch_tree<2>(h, *it))) return rc;
    }
    return MADIPM_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int32_t device_create(madipm_handle* h) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        h->error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
        return MADIPM_CUDA_ERROR;
    }
    if (h->cfg.device < 0 || h->cfg.device >= count) { h->error = "bad device ordinal"; return MADIPM_BAD_ARGUMENT; }
    CUDA_OK(h, cudaSetDevice(h->cfg.device));
    cudaDeviceProp prop;
    CUDA_OK(h, cudaGetDeviceProperties(&prop, h->cfg.device));
    if (prop.major < 10) { h->error = "device is not sm_100-class; this library targets B200 only"; return MADIPM_CUDA_ERROR; }

    DeviceState* d = new DeviceState();
    h->dev = d;
    d->sym = &h->sym;
    d->device = h->cfg.device;
    CUDA_OK(h, cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CUDA_OK(h, cudaEventCreate(&d->ev0));
    CUDA_OK(h, cudaEventCreate(&d->ev1));
    CUDA_OK(h, cudaEventCreate(&d->ev2));
    CUDA_OK(h, cudaEventCreate(&d->ev3));
    CUDA_OK(h, cudaEventCreate(&d->ev4));
    const KktPattern& p = h->pat;
    const Symbolic& s = h->sym;
    if (p.A_rowptr.back() > INT32_MAX || p.kkt_colptr.back() > INT32_MAX || (int64_t)s.sn_rows.size() > INT32_MAX ||
        (int64_t)p.prod_a.size() > INT32_MAX) {
        h->error = "pattern too large for 32-bit device indices";
        return MADIPM_BAD_ARGUMENT;
    }
    d->B = h->cfg.batch;
    d->n = p.n_tot; d->m = p.m; d->nlb = h->nlb; d->nub = h->nub; d->N = p.N; d->kind = p.kind;
    d->scaled = p.kind == 0 && h->cfg.kkt_kind == MADIPM_KKT_K2_SCALED;
    d->L = static_cast<int64_t>(p.n_tot) + p.m + h->nlb + h->nub;
    d->nnzA = p.A_rowptr.back(); d->nnzQ = p.Q_rowptr.empty() ? 0 : p.Q_rowptr.back();
    d->nnzh = p.nnzh; d->nnzj = p.nnzj; d->nnzJfull = p.nnzj + p.n_slack;
    d->nx = p.nx; d->n_slack = p.n_slack;
    d->kkt_nnz = p.kkt_colptr.back();
    d->lx_stride = s.nnz_stored;
    d->nsuper = s.nsuper; d->n_levels = s.n_levels;
    // Update groups of the big supernodes.  A wide fundamental supernode that was cut into pieces of kBigTile columns is a
    // chain s1 -> s2 -> ... whose row structures nest exactly (rows below s_q = columns of s_q+1 + rows below s_q+1).
    // Up to kBigGroup consecutive pieces form a group: a piece updates only the columns of the later pieces of its group
    // right away (upd_lim = that many tile columns), and the last piece applies the update of the WHOLE group to
    // everything beyond it in one pass (upd_nsrc sources, K = sum of their widths per read-modify-write of a target
    // tile instead of 64: the update is bound by that traffic).
    std::vector<int> upd_lim(s.big_sn.size(), -1), upd_nsrc(s.big_sn.size(), 1), upd_src(s.big_sn.size() * kBigGroup, 0);
    {
        std::vector<int> where(s.nsuper, -1);
        for (size_t q = 0; q < s.big_sn.size(); ++q) where[s.big_sn[q]] = static_cast<int>(q);
        auto height = [&](int sn) { return static_cast<int64_t>(s.sn_rowptr[sn + 1] - s.sn_rowptr[sn]); };
        auto width = [&](int sn) { return static_cast<int64_t>(s.sn_first[sn + 1] - s.sn_first[sn]); };
        const char* genv = std::getenv("MADIPM_BIG_GROUP");
        const int gmax = genv ? std::max(1, std::min(kBigGroup, std::atoi(genv))) : kBigGroup;
        std::vector<char> taken(s.nsuper, 0);
        for (size_t q = 0; q < s.big_sn.size(); ++q) {                     // big_sn is sorted by height
            const int first = s.big_sn[q];
            if (taken[first]) continue;
            std::vector<int> grp{first};
            taken[first] = 1;
            while (static_cast<int>(grp.size()) < gmax) {
                const int c = grp.back(), par = s.sn_parent[c];
                if (par < 0 || where[par] < 0 || taken[par] || width(c) != kBigTile || width(par) != kBigTile ||
                    height(c) - width(c) != height(par)) break;      // (tile columns must coincide with pieces)
                if (height(par) - width(par) < kBigGroupMinRows) break;   // small remainder: too few tiles to be worth deferring
                grp.push_back(par);
                taken[par] = 1;
            }
            const int ng = static_cast<int>(grp.size());
            for (int j = 0; j < ng; ++j) {
                const int qq = where[grp[j]];
                if (j + 1 < ng) {                                          // narrow: the columns of the later pieces
                    int64_t cols = 0;
                    for (int k = j + 1; k < ng; ++k) cols += width(grp[k]);
                    upd_lim[qq] = static_cast<int>((cols + kBigTile - 1) / kBigTile);
                    upd_nsrc[qq] = 1;
                    upd_src[qq * kBigGroup] = grp[j];
                } else {                                                   // wide: the whole group onto the rest
                    upd_lim[qq] = -1;
                    upd_nsrc[qq] = ng;
                    for (int k = 0; k < ng; ++k) upd_src[qq * kBigGroup + k] = grp[k];
                }
            }
        }
    }
    // grid sizes of the multi-block kernels, per height (max over the big supernodes of that height)
    d->big_ptr.assign(s.big_ptr.begin(), s.big_ptr.end());
    for (int l = 0; l < s.n_levels; ++l) {
        int rb = 0, rb512 = 0, tiles = 0, pb = 0;
        for (int q = s.big_ptr[l]; q < s.big_ptr[l + 1]; ++q) {
            const int sn = s.big_sn[q];
            const int64_t hs = s.sn_rowptr[sn + 1] - s.sn_rowptr[sn], ws = s.sn_first[sn + 1] - s.sn_first[sn];
            const int64_t nb = hs - ws, nt = (nb + kBigTile - 1) / kBigTile;
            rb = std::max<int>(rb, static_cast<int>(nt));
            rb512 = std::max<int>(rb512, static_cast<int>((nb + 511) / 512));
            int64_t ntile = nt * (nt + 1) / 2;
            if (upd_lim[q] >= 0) {
                ntile = 0;
                for (int64_t tj = 0; tj < std::min<int64_t>(upd_lim[q], nt); ++tj) ntile += nt - tj;
            }
            tiles = std::max<int>(tiles, static_cast<int>(ntile));
            if (ws < kPanelThreads) pb = std::max<int>(pb, static_cast<int>((nb + (kPanelThreads - ws) - 1) / (kPanelThreads - ws)));
        }
        d->big_panelblocks.push_back(pb);
        d->big_rowblocks.push_back(rb); d->big_rowblocks512.push_back(rb512); d->big_tiles.push_back(tiles);
        d->big_max_per_level = std::max(d->big_max_per_level, s.big_ptr[l + 1] - s.big_ptr[l]);
    }
    const int B = d->B;
    int32_t rc;
#define UP(T, dst, src) if ((rc = upload_as<T>(h, &d->dst, src))) return rc
    // index sets and inverse maps
    std::vector<int> lb_of(d->n, -1), ub_of(d->n, -1);
    for (int k = 0; k < h->nlb; ++k) lb_of[h->ind_lb[k]] = k;
    for (int k = 0; k < h->nub; ++k) ub_of[h->ind_ub[k]] = k;
    UP(int, ind_lb, h->ind_lb); UP(int, ind_ub, h->ind_ub); UP(int, lb_of, lb_of); UP(int, ub_of, ub_of);
    // operators
    UP(int, A_rp, p.A_rowptr); UP(int, A_ci, p.A_colind); UP(int, A_order, p.A_csr_order); UP(int, jac_row, p.jac_I);
    UP(int, AT_rp, p.AT_rowptr); UP(int, AT_ci, p.AT_colind); UP(int, AT_src, p.AT_src);
    std::vector<int64_t> qrp = p.Q_rowptr;
    if (qrp.empty()) qrp.assign(static_cast<size_t>(d->n) + 1, 0);
    UP(int, Q_rp, qrp); UP(int, Q_ci, p.Q_colind); UP(int, Q_src, p.Q_src);
    if ((rc = dev_alloc(h, &d->A_val, B * d->nnzA))) return rc;
    if ((rc = dev_alloc(h, &d->AT_val, B * d->nnzA))) return rc;
    if ((rc = dev_alloc(h, &d->Q_val, B * d->nnzQ))) return rc;
    if ((rc = dev_alloc(h, &d->hessV, B * d->nnzh))) return rc;
    if ((rc = dev_alloc(h, &d->jacV, B * d->nnzJfull))) return rc;
    // KKT matrix
    if (p.kind == 0) {
        UP(int, aug_map, p.aug_csc_map); UP(int, diag_slot, p.diag_slot);
        if ((rc = dev_alloc(h, &d->const_diag, static_cast<size_t>(B) * d->N))) return rc;
    } else {
        UP(int64_t, prod_ptr, p.prod_ptr); UP(int, prod_a, p.prod_a); UP(int, prod_b, p.prod_b); UP(int, prod_k, p.prod_k);
    }
    // factor structure
    UP(int, sn_first, s.sn_first); UP(int64_t, sn_rowptr, s.sn_rowptr); UP(int, sn_rows, s.sn_rows);
    UP(int64_t, sn_xoff, s.sn_xoff); UP(int, sn_of, s.sn_of); UP(int64_t, asm_map, s.asm_map);
    UP(int64_t, row_colbase, s.row_colbase); UP(int64_t, row_relbase, s.row_relbase); UP(int, rel, s.rel);
    UP(int, perm, s.perm); UP(int, iperm, s.iperm);
    UP(int, big_sn, s.big_sn);
    UP(int, upd_lim, upd_lim); UP(int, upd_nsrc, upd_nsrc); UP(int, upd_src, upd_src);
    UP(int, par_pos, s.par_pos);
    {
        std::vector<SnMeta> meta(std::max<size_t>(1, s.task_sn.size()));      // in task order
        std::vector<ChildRec> crec(std::max<size_t>(1, s.child_idx.size()));
        std::memset(meta.data(), 0, meta.size() * sizeof(SnMeta));
        std::memset(crec.data(), 0, crec.size() * sizeof(ChildRec));
        for (size_t q = 0; q < s.task_sn.size(); ++q) {
            const int sn = s.task_sn[q];
            SnMeta& m = meta[q];
            m.xoff = s.sn_xoff[sn]; m.coff = s.sn_coff[sn]; m.rowptr = static_cast<int32_t>(s.sn_rowptr[sn]);
            m.fc = s.sn_first[sn]; m.h = static_cast<int32_t>(s.sn_rowptr[sn + 1] - s.sn_rowptr[sn]);
            m.w = s.sn_first[sn + 1] - s.sn_first[sn]; m.kind = s.sn_kind[sn]; m.sn = sn;
            m.c0 = s.child_ptr[sn]; m.c1 = s.child_ptr[sn + 1];
        }
        for (int sn = 0; sn < s.nsuper; ++sn) {
            const int wp = s.sn_first[sn + 1] - s.sn_first[sn];
            for (int q = s.child_ptr[sn]; q < s.child_ptr[sn + 1]; ++q) {
                const int c = s.child_idx[q];
                const int wc = s.sn_first[c + 1] - s.sn_first[c];
                ChildRec& r = crec[q];
                r.coff = s.sn_coff[c];
                r.below = static_cast<int32_t>(s.sn_rowptr[c] + wc);
                r.nb = static_cast<int32_t>(s.sn_rowptr[c + 1] - s.sn_rowptr[c]) - wc;
                r.pull = s.sn_kind[c] == 0;
                r.nbp = 0;
                while (r.nbp < r.nb && s.par_pos[r.below + r.nbp] < wp) ++r.nbp;
            }
        }
        if (s.max_w > 65535) { h->error = "max_width too large"; return MADIPM_BAD_ARGUMENT; }
        if ((rc = dev_alloc(h, &d->sn_meta, meta.size()))) return rc;
        CUDA_OK(h, cudaMemcpy(d->sn_meta, meta.data(), meta.size() * sizeof(SnMeta), cudaMemcpyHostToDevice));
        if ((rc = dev_alloc(h, &d->child_rec, crec.size()))) return rc;
        CUDA_OK(h, cudaMemcpy(d->child_rec, crec.data(), crec.size() * sizeof(ChildRec), cudaMemcpyHostToDevice));
        // dataflow scheduling: the task of every supernode's parent, one dependency counter per (instance, task)
        d->ntasks = static_cast<int>(s.task_sn.size());
        std::vector<int> sn_parent(std::max(1, s.nsuper), -1), task_of(std::max(1, s.nsuper), -1), tpar(std::max<size_t>(1, s.task_sn.size()), -1);
        for (int sn = 0; sn < s.nsuper; ++sn)
            for (int q = s.child_ptr[sn]; q < s.child_ptr[sn + 1]; ++q) sn_parent[s.child_idx[q]] = sn;
        for (size_t q = 0; q < s.task_sn.size(); ++q) task_of[s.task_sn[q]] = static_cast<int>(q);
        for (size_t q = 0; q < s.task_sn.size(); ++q) {
            const int ps = sn_parent[s.task_sn[q]];
            tpar[q] = ps >= 0 ? task_of[ps] : -1;
        }
        if ((rc = dev_alloc(h, &d->task_parent, tpar.size()))) return rc;
        CUDA_OK(h, cudaMemcpy(d->task_parent, tpar.data(), tpar.size() * sizeof(int), cudaMemcpyHostToDevice));
        if ((rc = dev_alloc(h, &d->dep, static_cast<size_t>(B) * std::max(1, d->ntasks)))) return rc;
        d->phase_ptr.assign(s.phase_ptr.begin(), s.phase_ptr.end());
    }
    {
        std::vector<int32_t> pc = s.phase_cls;
        if (pc.empty()) pc.assign(5, 0);
        UP(int, phase_cls, pc);
        for (int op = 0; op < 3; ++op) {
            if ((rc = dev_alloc(h, &d->tickets[op], static_cast<size_t>(kSyncStride) * std::max(1, s.n_levels)))) return rc;
            CUDA_OK(h, cudaMemsetAsync(d->tickets[op], 0, sizeof(int) * kSyncStride * std::max(1, s.n_levels), d->stream));
        }
    }
    std::vector<int64_t> diag_pos(d->N);
    for (int sn = 0; sn < s.nsuper; ++sn) {
        int64_t hs = s.sn_rowptr[sn + 1] - s.sn_rowptr[sn];
        for (int j = s.sn_first[sn]; j < s.sn_first[sn + 1]; ++j)
            diag_pos[j] = s.sn_xoff[sn] + static_cast<int64_t>(j - s.sn_first[sn]) * (hs + 1);
    }
    UP(int64_t, diag_pos, diag_pos);
    if (d->scaled) {                                       // no panel image: every entry moves every iteration
        std::vector<int> kcol(p.kkt_colptr.back());
        for (int j = 0; j < d->N; ++j)
            for (int64_t k = p.kkt_colptr[j]; k < p.kkt_colptr[j + 1]; ++k) kcol[k] = j;
        UP(int, kkt_col, kcol);
        UP(int, kkt_row, p.kkt_rowval);
        if ((rc = dev_alloc(h, &d->nz_base, static_cast<size_t>(h->cfg.batch) * p.kkt_colptr.back()))) return rc;
    } else if (p.kind == 0 && s.n_big == 0) {
        std::vector<int64_t> diag_lx(d->N);
        for (int i = 0; i < d->N; ++i) diag_lx[i] = s.asm_map[p.diag_slot[i]];
        UP(int64_t, diag_lx, diag_lx);
        if ((rc = dev_alloc(h, &d->P0, static_cast<size_t>(h->cfg.batch) * s.nnz_stored))) return rc;
        if ((rc = dev_alloc(h, &d->dperm, static_cast<size_t>(h->cfg.batch) * d->N))) return rc;
    }
#undef UP
    if ((rc = dev_alloc(h, &d->Lx, static_cast<size_t>(B) * d->lx_stride))) return rc;
    if ((rc = dev_alloc(h, &d->xw, static_cast<size_t>(B) * d->N))) return rc;
    if ((rc = dev_alloc(h, &d->fail, B))) return rc;
    d->h_fail.assign(B, 0);
    // contribution blocks, contribution vectors, grid barrier
    d->cx_stride = s.sn_coff.empty() ? 0 : s.sn_coff.back();
    d->cv_stride = static_cast<int64_t>(s.sn_rows.size());
    if ((rc = dev_alloc(h, &d->Cx, static_cast<size_t>(B) * d->cx_stride + 4))) return rc;
    if ((rc = dev_alloc(h, &d->cvec, static_cast<size_t>(B) * d->cv_stride))) return rc;
    if ((rc = dev_alloc(h, &d->abort_flag, 1))) return rc;
    CUDA_OK(h, cudaMemsetAsync(d->abort_flag, 0, sizeof(unsigned), d->stream));
    CUDA_OK(h, cudaHostAlloc(reinterpret_cast<void**>(&d->tree_dbg), sizeof(TreeDebug), cudaHostAllocMapped | cudaHostAllocPortable));
    std::memset(d->tree_dbg, 0, sizeof(TreeDebug));
    if (d->kind != 0) d->tiny_pivot = d->growth_limit = 0.0;   // normal equations: pivots and factor entries legitimately span 1e-12 .. 1e12 (Sigma does)
    if (const char* e = std::getenv("MADIPM_SYNC_TIMEOUT_S")) d->sync_timeout_s = std::max(0.0, std::atof(e));
    if (const char* e = std::getenv("MADIPM_BARRIER_TIMEOUT_MS")) d->spin_budget_ns = std::max(1ll, std::atoll(e)) * 1000000ll;
    d->num_sms = prop.multiProcessorCount;
    {
        // shared memory of the tree kernels: one 256-thread group or 2 / 4 / 8 smaller ones (sizes: symbolic.cpp step 8)
        d->tree_maxw = std::max(s.max_w, 1);
        d->smem_tree = static_cast<size_t>(s.tree_slice[0]) * sizeof(double);
        if (d->smem_tree > 220 * 1024) { h->error = "block_panel too large for the shared memory of one block"; return MADIPM_BAD_ARGUMENT; }
        if (!prop.cooperativeLaunch) { h->error = "device does not support cooperative launches"; return MADIPM_CUDA_ERROR; }
        CUDA_OK(h, cudaFuncSetAttribute(k_tree<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree_top<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree_top<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree_top<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree_top<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->tree_occ[0], k_tree<0>, kTreeThreads, d->smem_tree));
        CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->tree_occ[1], k_tree<1>, kTreeThreads, d->smem_tree));
        CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->tree_occ[2], k_tree<2>, kTreeThreads, d->smem_tree));
        CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->tree_occ[3], k_tree<3>, kTreeThreads, d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree_df<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree_df<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree_df<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaFuncSetAttribute(k_tree_df<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_tree));
        CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->tree_occ_df[0], k_tree_df<0>, kTreeThreads, d->smem_tree));
        CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->tree_occ_df[1], k_tree_df<1>, kTreeThreads, d->smem_tree));
        CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->tree_occ_df[2], k_tree_df<2>, kTreeThreads, d->smem_tree));
        CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d->tree_occ_df[3], k_tree_df<3>, kTreeThreads, d->smem_tree));
        int top_occ = 1;
        {
            int o = 0;
            CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_tree_top<0>, kTopThreads, d->smem_tree)); top_occ = std::min(top_occ, o);
            CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_tree_top<1>, kTopThreads, d->smem_tree)); top_occ = std::min(top_occ, o);
            CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_tree_top<2>, kTopThreads, d->smem_tree)); top_occ = std::min(top_occ, o);
            CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k_tree_top<3>, kTopThreads, d->smem_tree)); top_occ = std::min(top_occ, o);
        }
        {   // k_front for the factorisation of the heights near the root: needs every front to fit its layout
            auto even = [](int64_t v) { return static_cast<int>((v + 1) & ~int64_t(1)); };
            int64_t nbmax = 0;
            for (int sn = 0; sn < s.nsuper; ++sn)
                if (s.sn_kind[sn] != 2)
                    nbmax = std::max<int64_t>(nbmax, (s.sn_rowptr[sn + 1] - s.sn_rowptr[sn]) - (s.sn_first[sn + 1] - s.sn_first[sn]));
            FrontLayout lo;
            lo.cap = even(std::max(s.small_max_panel, 2)) + 2;
            lo.ctri = even(nbmax * (nbmax + 1) / 2 + 2);
            lo.maxh = even(std::max(s.small_max_h, 2));
            lo.maxw = even(std::max(s.max_w, 1));
            lo.total = lo.cap + lo.ctri + lo.maxh + lo.maxw + 2 * 66 + kChildChunk * 4 + kFrontPosCap / 2;
            const char* nf = std::getenv("MADIPM_NO_FRONT");
            d->front_lo = lo;
            d->smem_front = static_cast<size_t>(lo.total) * sizeof(double);
            d->use_front = s.n_big == 0 && s.small_max_h <= kFrontThreads && s.max_w <= 64 && d->smem_front <= 200 * 1024 &&
                           !(nf && nf[0] == '1');
            if (d->use_front) {
                CUDA_OK(h, cudaFuncSetAttribute(k_front<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_front));
                CUDA_OK(h, cudaFuncSetAttribute(k_front<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_front));
                int o0 = 0, o3 = 0;
                CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o0, k_front<0>, kFrontThreads, d->smem_front));
                CUDA_OK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, k_front<3>, kFrontThreads, d->smem_front));
                if (o0 < 1 || o3 < 1) d->use_front = false;
            }
        }
        for (int op = 0; op < 4; ++op)
            if (d->tree_occ[op] < 1) { h->error = "tree kernel does not fit on an SM"; return MADIPM_INTERNAL; }
        // launch plan: consecutive heights share one cooperative launch; a height with big supernodes ends it
        // (MADIPM_TREE_SPLIT=1 gives every height its own launch: a per-height launch list for the profiler;
        //  MADIPM_TREE_NOTOP=1 keeps the heights near the root in k_tree)
        const char* split_env = std::getenv("MADIPM_TREE_SPLIT");
        const bool split = split_env && split_env[0] == '1';
        const char* notop_env = std::getenv("MADIPM_TREE_NOTOP");
        const bool use_top = top_occ >= 1 && !(notop_env && notop_env[0] == '1');
        int l0 = 0;
        for (int l = 0; l < s.n_levels; ++l) {
            const bool last = l + 1 == s.n_levels;
            if (s.big_ptr[l + 1] > s.big_ptr[l] || last || split) {
                DeviceState::TreeSeg sg;
                sg.ph0 = l0; sg.ph1 = l + 1; sg.split = sg.ph1;
                auto ntasks = [&](int q) { return static_cast<int64_t>(s.phase_ptr[q + 1] - s.phase_ptr[q]) * B; };
                if (use_top)
                    while (sg.split > sg.ph0 && ntasks(sg.split - 1) <= d->num_sms) --sg.split;
                int64_t need = 0, need_top = 0;
                for (int q = sg.ph0; q < sg.split; ++q) {
                    const int32_t* pc = s.phase_cls.data() + static_cast<size_t>(q) * 5;
                    int64_t blocks = 0;
                    for (int c = 0; c < 4; ++c) blocks += (static_cast<int64_t>(pc[c + 1] - pc[c]) * B + (1 << c) - 1) >> c;
                    need = std::max(need, blocks);
                }
                for (int q = sg.split; q < sg.ph1; ++q) need_top = std::max(need_top, ntasks(q));
                for (int op = 0; op < 4; ++op)
                    sg.grid[op] = static_cast<int>(std::min<int64_t>(need, static_cast<int64_t>(d->tree_occ[op]) * d->num_sms));
                sg.grid_top = static_cast<int>(std::min<int64_t>(need_top, d->num_sms));
                // Dataflow for the triangular sweeps only where some height of the launch is short of tasks (its barrier then
                // costs the latency of a front): with >= 2 tasks per block at EVERY height the sweeps are throughput-bound and
                // the per-task signalling costs more than the barriers (measured: C5 with 4096 instances 0.86 -> 0.97 ms per
                // solve, with 1024 instances 0.31 -> 0.29 ms, C3 0.70 -> 0.65 ms; 2048 instances were not measured and stay with the
                // barriers: the threshold sits just above the 1024-instance case on 592 blocks).  The factorisation gains in every case.
                int64_t min_tasks = INT64_MAX;
                for (int q = sg.ph0; q < sg.split; ++q) min_tasks = std::min(min_tasks, ntasks(q));
                sg.df_sweeps = min_tasks < 2 * static_cast<int64_t>(std::max(1, sg.grid[1]));
                d->segs.push_back(sg);
                l0 = l + 1;
            }
        }
        // dataflow scheduling needs the whole small tree below the split in ONE launch that starts at the leaves
        d->dataflow_ok = s.n_big == 0 && !split && d->segs.size() == 1 && d->segs[0].ph0 == 0 && d->ntasks > 0;
        for (int op = 0; op < 4; ++op) d->dataflow_ok = d->dataflow_ok && d->tree_occ_df[op] >= 1;
        d->dataflow = true;                               // (measured faster on C3 and on C5 shards; bit-identical results)
        if (const char* e = std::getenv("MADIPM_TREE_DATAFLOW")) d->dataflow = e[0] != '0';
        d->tri_cap = std::min(s.max_w * s.max_w, 4096);
        d->smem_tri = (static_cast<size_t>(d->tri_cap) + 256) * sizeof(double);
    }
    {
        const size_t mw = static_cast<size_t>(s.max_w);
        d->smem_big_diag = (2 * mw * mw + mw) * sizeof(double);
        d->smem_big_trsm = 2 * mw * kBigTileLd * sizeof(double);
        d->smem_big_update = (2 * kBigStage + kBigGroup * std::max<size_t>(kBigTile, (mw + 1) & ~size_t(1))) * sizeof(double) + 2 * kBigTile * sizeof(int64_t);
        if (!s.big_sn.empty()) {
            if (d->smem_big_trsm > 220 * 1024) { h->error = "max_width too large for the big-front kernels"; return MADIPM_BAD_ARGUMENT; }
            CUDA_OK(h, cudaFuncSetAttribute(k_big_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_big_diag));
            CUDA_OK(h, cudaFuncSetAttribute(k_big_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_big_trsm));
            CUDA_OK(h, cudaFuncSetAttribute(k_big_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_big_update));
            d->big_use_panel = s.max_w <= 64;                 // one-launch panel factorisation (k_big_panel)
            d->smem_big_panel = static_cast<size_t>(kPanelThreads) * mw * sizeof(double) + 64;
            if (d->big_use_panel)
                CUDA_OK(h, cudaFuncSetAttribute(k_big_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->smem_big_panel));
            const size_t nbuf = static_cast<size_t>(B) * d->big_max_per_level * 256;
            if ((rc = dev_alloc(h, &d->big_done, static_cast<size_t>(B) * d->big_max_per_level))) return rc;
            CUDA_OK(h, cudaMemsetAsync(d->big_done, 0, static_cast<size_t>(B) * d->big_max_per_level * sizeof(int), d->stream));
            if ((rc = dev_alloc(h, &d->big_tbuf, nbuf))) return rc;
            if ((rc = dev_alloc(h, &d->big_zbuf, static_cast<size_t>(B) * d->big_max_per_level * mw * mw))) return rc;
            d->big_zstride = static_cast<int>(mw * mw);
            CUDA_OK(h, cudaMemsetAsync(d->big_tbuf, 0, nbuf * sizeof(double), d->stream));
        }
    }

    // vectors
    const int64_t n = d->n, m = d->m;
    int64_t* vl = d->veclen;
    for (int v : {MADIPM_VEC_X, MADIPM_VEC_XL, MADIPM_VEC_XU, MADIPM_VEC_ZL, MADIPM_VEC_ZU, MADIPM_VEC_F, MADIPM_VEC_JACL,
                  MADIPM_VEC_COST, MADIPM_VEC_REG, MADIPM_VEC_PR_DIAG, MADIPM_VEC_SCALING}) vl[v] = n;
    for (int v : {MADIPM_VEC_Y, MADIPM_VEC_C, MADIPM_VEC_RHS, MADIPM_VEC_DU_DIAG, MADIPM_VEC_CON_SCALE}) vl[v] = m;
    for (int v : {MADIPM_VEC_D, MADIPM_VEC_P, MADIPM_VEC_W1, MADIPM_VEC_W2}) vl[v] = d->L;
    vl[MADIPM_VEC_CORR_LB] = vl[MADIPM_VEC_L_DIAG] = vl[MADIPM_VEC_L_LOWER] = d->nlb;
    vl[MADIPM_VEC_CORR_UB] = vl[MADIPM_VEC_U_DIAG] = vl[MADIPM_VEC_U_LOWER] = d->nub;
    vl[MADIPM_VEC_KKT_NZVAL] = d->kkt_nnz;
    for (int v = 0; v < MADIPM_VEC_COUNT; ++v) {
        if ((rc = dev_alloc(h, &d->vec[v], static_cast<size_t>(B) * vl[v]))) return rc;
        CUDA_OK(h, cudaMemsetAsync(d->vec[v], 0, static_cast<size_t>(B) * std::max<int64_t>(vl[v], 1) * sizeof(double), d->stream));
    }
    if ((rc = dev_alloc(h, &d->buffer_n, static_cast<size_t>(B) * n))) return rc;
    if ((rc = dev_alloc(h, &d->buffer_m, static_cast<size_t>(B) * m))) return rc;
    // reductions
    if ((rc = dev_alloc(h, &d->red_scratch, static_cast<size_t>(B) * kMaxRedBlocks * 8))) return rc;
    if ((rc = dev_alloc(h, &d->red_counter, B))) return rc;
    CUDA_OK(h, cudaMemsetAsync(d->red_counter, 0, B * sizeof(unsigned int), d->stream));
    if ((rc = dev_alloc(h, &d->d_out, static_cast<size_t>(B) * 16))) return rc;
    if ((rc = dev_alloc(h, &d->d_out2, static_cast<size_t>(B) * 16))) return rc;
    if ((rc = dev_alloc(h, &d->d_norms, static_cast<size_t>(B) * 4))) return rc;
    if ((rc = dev_alloc(h, &d->d_obj, static_cast<size_t>(B) * 2))) return rc;
    if ((rc = dev_alloc(h, &d->d_iter, static_cast<size_t>(B) * 19))) return rc;
    if ((rc = dev_alloc(h, &d->d_ok, 1))) return rc;
    CUDA_OK(h, cudaMallocHost(reinterpret_cast<void**>(&d->h_iter), (static_cast<size_t>(B) * 32 + 2) * sizeof(double)));
    CUDA_OK(h, cudaMallocHost(reinterpret_cast<void**>(&d->h_fail_pinned), static_cast<size_t>(B) * sizeof(int)));
    if ((rc = dev_alloc(h, &d->d_steps, static_cast<size_t>(B) * 2))) return rc;
    if ((rc = dev_alloc(h, &d->d_mu, B))) return rc;
    if ((rc = dev_alloc(h, &d->d_tau, B))) return rc;
    if ((rc = dev_alloc(h, &d->d_objscale, B))) return rc;
    if ((rc = dev_alloc(h, &d->d_ones, B))) return rc;
    if ((rc = dev_alloc(h, &d->d_active, B))) return rc;
    if ((rc = dev_alloc(h, &d->d_inertia, static_cast<size_t>(B) * 3))) return rc;
    LAUNCH(h, k_fill, dim3(1, 1), 256, 0, (int64_t)B, d->d_ones, (int64_t)0, 1.0);
    CUDA_OK(h, cudaMallocHost(reinterpret_cast<void**>(&d->h_out), static_cast<size_t>(B) * 32 * sizeof(double)));
    { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }
    return MADIPM_OK;
}

void device_destroy(madipm_handle* h) {
    DeviceState* d = h->dev;
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->stream && !d->wedged) sync_stream(h);
    if (d->wedged) {                  // a kernel that never returns owns the stream: cudaFree would block for ever -- leak instead
        delete d;
        h->dev = nullptr;
        return;
    }
    for (void* p : d->allocs) cudaFree(p);
    if (d->h_out) cudaFreeHost(d->h_out);
    if (d->h_fail_pinned) cudaFreeHost(d->h_fail_pinned);
    if (d->h_iter) cudaFreeHost(d->h_iter);
    if (d->tree_dbg) cudaFreeHost(d->tree_dbg);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->ev2) cudaEventDestroy(d->ev2);
    if (d->ev3) cudaEventDestroy(d->ev3);
    if (d->ev4) cudaEventDestroy(d->ev4);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
    h->dev = nullptr;
}

}  // namespace madipm

using namespace madipm;

extern "C" {

// ---- values -----------------------------------------------------------------------------------
// host -> device copy of the coordinate values (slack columns of the Jacobian: -1, normalkkt.jl:151)
static int32_t upload_values(madipm_handle* h, const double* hess_V, const double* jac_V) {
    DeviceState* d = h->dev;
    const int B = d->B;
    if ((d->nnzh > 0 && !hess_V) || (d->nnzj > 0 && !jac_V)) return MADIPM_BAD_ARGUMENT;
    { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }      // pageable sources below: wait for the stream under the guard
    if (d->nnzh) CUDA_OK(h, cudaMemcpyAsync(d->hessV, hess_V, sizeof(double) * B * d->nnzh, cudaMemcpyHostToDevice, d->stream));
    if (d->nnzJfull > d->nnzj)
        LAUNCH(h, k_fill, dim3(vec_blocks(d->nnzJfull), B), kVecThreads, 0, d->nnzJfull, d->jacV, d->nnzJfull, -1.0);
    if (d->nnzj)
        CUDA_OK(h, cudaMemcpy2DAsync(d->jacV, sizeof(double) * d->nnzJfull, jac_V, sizeof(double) * d->nnzj,
                                     sizeof(double) * d->nnzj, B, cudaMemcpyHostToDevice, d->stream));
    return MADIPM_OK;
}

// compress_jacobian! / compress_hessian! from the device copies of the coordinate values
static int32_t compress_values(madipm_handle* h) {
    DeviceState* d = h->dev;
    const int B = d->B;
    // compress_jacobian! (normalkkt.jl:149-158): AT.nzval[i] = A.V[A_csr_map[i]]
    if (d->nnzA) {
        LAUNCH(h, k_gather, dim3(vec_blocks(d->nnzA), B), kVecThreads, 0, d->nnzA, d->A_order, d->jacV, d->nnzJfull, d->A_val, d->nnzA);
        LAUNCH(h, k_gather, dim3(vec_blocks(d->nnzA), B), kVecThreads, 0, d->nnzA, d->AT_src, d->A_val, d->nnzA, d->AT_val, d->nnzA);
    }
    if (d->nnzQ)
        LAUNCH(h, k_gather, dim3(vec_blocks(d->nnzQ), B), kVecThreads, 0, d->nnzQ, d->Q_src, d->hessV, d->nnzh, d->Q_val, d->nnzQ);
    if (d->kind == 0) {
        // constant part of aug_com.nzval: Hessian and Jacobian blocks (duplicates summed); diagonals later
        double* nz = d->vec[MADIPM_VEC_KKT_NZVAL];
        CUDA_OK(h, cudaMemsetAsync(nz, 0, sizeof(double) * B * d->kkt_nnz, d->stream));
        if (d->nnzh)
            LAUNCH(h, k_scatter_add, dim3(vec_blocks(d->nnzh), B), kVecThreads, 0, d->nnzh, d->aug_map + d->n, d->hessV, d->nnzh, nz, d->kkt_nnz);
        if (d->nnzJfull)
            LAUNCH(h, k_scatter_add, dim3(vec_blocks(d->nnzJfull), B), kVecThreads, 0, d->nnzJfull, d->aug_map + d->n + d->nnzh, d->jacV, d->nnzJfull, nz, d->kkt_nnz);
        LAUNCH(h, k_extract_diag, dim3(vec_blocks(d->N), B), kVecThreads, 0, d->N, d->diag_slot, nz, d->kkt_nnz, d->const_diag);
        if (d->scaled)
            CUDA_OK(h, cudaMemcpyAsync(d->nz_base, nz, sizeof(double) * B * d->kkt_nnz, cudaMemcpyDeviceToDevice, d->stream));
        if (d->P0) {                                       // panel image of the constant part
            CUDA_OK(h, cudaMemsetAsync(d->P0, 0, sizeof(double) * B * d->lx_stride, d->stream));
            LAUNCH(h, k_assemble_panels, dim3(vec_blocks(d->kkt_nnz) * 4, B), kVecThreads, 0, d->kkt_nnz, d->asm_map, nz, d->kkt_nnz, d->P0,
                   d->lx_stride);
            d->image_valid = true;
        }
    }
    d->values_set = true;
    return MADIPM_OK;
}

int32_t madipm_set_values(madipm_handle* h, const double* hess_V, const double* jac_V) {
    int32_t rc = check(h);
    if (rc) return rc;
    Timer t(h, MADIPM_T_ASSEMBLY);
    if ((rc = upload_values(h, hess_V, jac_V))) return rc;
    return compress_values(h);
}

int32_t madipm_load_problem(madipm_handle* h, const double* x0, const double* lvar, const double* uvar, const double* y0,
                            const double* lcon, const double* ucon, const double* c, const double* hess_V, const double* jac_V,
                            const madipm_load_options* opt, double* obj_scale, double* norm_b) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    const int B = d->B, nx = d->nx, n = d->n, m = d->m;
    if (!opt || !obj_scale || !norm_b || (nx > 0 && (!x0 || !lvar || !uvar || !c)) || (m > 0 && (!y0 || !lcon || !ucon)))
        return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_ASSEMBLY);
    if ((rc = sync_stream(h))) return rc;                   // pageable sources below: wait for the stream under the guard
    auto up_n = [&](int vec, const double* src) {           // [B][nx] host -> the first nx entries of an n_tot vector
        return nx == 0 ? cudaSuccess
                       : cudaMemcpy2DAsync(d->vec[vec], sizeof(double) * n, src, sizeof(double) * nx, sizeof(double) * nx, B,
                                           cudaMemcpyHostToDevice, d->stream);
    };
    auto up_m = [&](int vec, const double* src) {
        return m == 0 ? cudaSuccess : cudaMemcpyAsync(d->vec[vec], src, sizeof(double) * B * m, cudaMemcpyHostToDevice, d->stream);
    };
    CUDA_OK(h, up_n(MADIPM_VEC_X, x0)); CUDA_OK(h, up_n(MADIPM_VEC_XL, lvar)); CUDA_OK(h, up_n(MADIPM_VEC_XU, uvar));
    CUDA_OK(h, up_n(MADIPM_VEC_COST, c));
    CUDA_OK(h, up_m(MADIPM_VEC_Y, y0)); CUDA_OK(h, up_m(MADIPM_VEC_RHS, lcon)); CUDA_OK(h, up_m(MADIPM_VEC_C, ucon));
    if ((rc = upload_values(h, hess_V, jac_V))) return rc;
    VecView V = make_view(*d);
    const dim3 gn(vec_blocks(n), B), gm(vec_blocks(m), B);
    // initialize!: relaxed bounds, start pushed inside, slacks at the constraint values of the pushed start
    LAUNCH(h, k_load_variables, gn, kVecThreads, 0, V, nx, opt->tol, opt->bound_push, opt->bound_fac);
    if (d->nnzA)
        LAUNCH(h, k_gather, dim3(vec_blocks(d->nnzA), B), kVecThreads, 0, d->nnzA, d->A_order, d->jacV, d->nnzJfull, d->A_val, d->nnzA);
    if (d->n_slack > 0) {
        spmv_launch(h, m, d->nnzA, d->A_rp, d->A_ci, d->A_val, d->nnzA, d->vec[MADIPM_VEC_X], n, d->buffer_m, m, 1.0, 0.0);
        LAUNCH(h, k_load_slacks, dim3(vec_blocks(d->n_slack), B), kVecThreads, 0, V, nx, d->n_slack, d->jac_row + d->nnzj, d->buffer_m,
               (int64_t)m, opt->tol, opt->bound_push, opt->bound_fac);
    }
    // set_scaling!
    double* cs = d->vec[MADIPM_VEC_CON_SCALE];
    std::vector<double> os(B, 1.0);
    if (opt->scaling) {
        if (m > 0) LAUNCH(h, k_con_scale, gm, kVecThreads, 0, m, d->A_rp, d->A_val, d->nnzA, opt->max_gradient, cs, (int64_t)m);
        if (nx > 0) {
            if (d->nnzQ)
                LAUNCH(h, k_gather, dim3(vec_blocks(d->nnzQ), B), kVecThreads, 0, d->nnzQ, d->Q_src, d->hessV, d->nnzh, d->Q_val, d->nnzQ);
            LAUNCH(h, k_grad0_max, dim3(vec_blocks(nx), B), kVecThreads, 0, V, nx, d->Q_rp, d->Q_ci, d->Q_val, d->nnzQ, d->nnzQ > 0 ? 1 : 0,
                   d->red_scratch, d->red_counter, d->d_out);
            if ((rc = fetch(h, d->d_out, os.data(), B))) return rc;
            for (int b = 0; b < B; ++b) os[b] = os[b] > 0.0 ? std::min(1.0, opt->max_gradient / os[b]) : 1.0;
        }
    } else if (m > 0) {
        LAUNCH(h, k_fill, gm, kVecThreads, 0, (int64_t)m, cs, (int64_t)m, 1.0);
    }
    CUDA_OK(h, cudaMemcpyAsync(d->d_objscale, os.data(), sizeof(double) * B, cudaMemcpyHostToDevice, d->stream));
    LAUNCH(h, k_load_scale, dim3(vec_blocks(std::max(n, m)), B), kVecThreads, 0, V, nx, d->n_slack, d->jac_row + d->nnzj, cs, (int64_t)m,
           d->d_objscale);
    if (opt->scaling && (d->nnzh > 0 || d->nnzj > 0))
        LAUNCH(h, k_scale_values, dim3(vec_blocks(std::max(d->nnzh, d->nnzj)), B), kVecThreads, 0, d->nnzh, d->hessV, d->nnzh, d->nnzj,
               d->jacV, d->nnzJfull, d->jac_row, cs, (int64_t)m, d->d_objscale);
    if ((rc = compress_values(h))) return rc;
    for (int b = 0; b < B; ++b) norm_b[b] = 0.0;
    if (m > 0) {
        LAUNCH(h, k_norm_inf, gm, kVecThreads, 0, (int64_t)m, d->vec[MADIPM_VEC_RHS], (int64_t)m, d->red_scratch, d->red_counter, d->d_out);
        if ((rc = fetch(h, d->d_out, norm_b, B))) return rc;
    } else {
        { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }       // `os` is read by the copy above
    }
    for (int b = 0; b < B; ++b) obj_scale[b] = os[b];
    return MADIPM_OK;
}

// ---- vectors ----------------------------------------------------------------------------------
int32_t madipm_vec_len(const madipm_handle* h, int32_t vec, int64_t* len) {
    if (!h || !len || vec < 0 || vec >= MADIPM_VEC_COUNT) return MADIPM_BAD_ARGUMENT;
    const KktPattern& p = h->pat;
    int64_t n = p.n_tot, m = p.m, L = n + m + h->nlb + h->nub;
    switch (vec) {
        case MADIPM_VEC_Y: case MADIPM_VEC_C: case MADIPM_VEC_RHS: case MADIPM_VEC_DU_DIAG: case MADIPM_VEC_CON_SCALE: *len = m; break;
        case MADIPM_VEC_D: case MADIPM_VEC_P: case MADIPM_VEC_W1: case MADIPM_VEC_W2: *len = L; break;
        case MADIPM_VEC_CORR_LB: case MADIPM_VEC_L_DIAG: case MADIPM_VEC_L_LOWER: *len = h->nlb; break;
        case MADIPM_VEC_CORR_UB: case MADIPM_VEC_U_DIAG: case MADIPM_VEC_U_LOWER: *len = h->nub; break;
        case MADIPM_VEC_KKT_NZVAL: *len = p.kkt_colptr.empty() ? 0 : p.kkt_colptr.back(); break;
        default: *len = n;
    }
    return MADIPM_OK;
}

// KKT_NZVAL is about to be read (or handed out): write the diagonal slots build_kkt skipped while the panel image was in use
static int32_t refresh_nzval_diag(madipm_handle* h) {
    DeviceState* d = h->dev;
    if (!d->nzval_diag_stale) return MADIPM_OK;
    LAUNCH(h, k_build_kkt_k2, dim3(vec_blocks(d->N), d->B), kVecThreads, 0, d->n, d->m, d->diag_slot, d->const_diag,
           d->vec[MADIPM_VEC_PR_DIAG], d->vec[MADIPM_VEC_DU_DIAG], d->vec[MADIPM_VEC_KKT_NZVAL], d->kkt_nnz);
    d->nzval_diag_stale = false;
    return MADIPM_OK;
}

int32_t madipm_vec_set(madipm_handle* h, int32_t vec, const double* host) {
    int32_t rc = check(h);
    if (rc) return rc;
    if (vec < 0 || vec >= MADIPM_VEC_COUNT || !host) return MADIPM_BAD_ARGUMENT;
    DeviceState* d = h->dev;
    if (vec == MADIPM_VEC_KKT_NZVAL) { d->image_valid = false; d->nzval_diag_stale = false; }   // the caller owns the KKT values now: assemble from nzval
    if ((rc = sync_stream(h))) return rc;      // (a copy from pageable memory synchronises the stream in the driver: do it under the guard)
    CUDA_OK(h, cudaMemcpyAsync(d->vec[vec], host, sizeof(double) * d->B * d->veclen[vec], cudaMemcpyHostToDevice, d->stream));
    return MADIPM_OK;
}

int32_t madipm_vec_get(madipm_handle* h, int32_t vec, double* host) {
    int32_t rc = check(h);
    if (rc) return rc;
    if (vec < 0 || vec >= MADIPM_VEC_COUNT || !host) return MADIPM_BAD_ARGUMENT;
    DeviceState* d = h->dev;
    if (vec == MADIPM_VEC_KKT_NZVAL && (rc = refresh_nzval_diag(h))) return rc;
    if ((rc = sync_stream(h))) return rc;      // a copy to pageable memory blocks in the driver: wait for the stream under the guard first
    CUDA_OK(h, cudaMemcpyAsync(host, d->vec[vec], sizeof(double) * d->B * d->veclen[vec], cudaMemcpyDeviceToHost, d->stream));
    { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }
    return MADIPM_OK;
}

int32_t madipm_vec_fill(madipm_handle* h, int32_t vec, double value) {
    int32_t rc = check(h);
    if (rc) return rc;
    if (vec < 0 || vec >= MADIPM_VEC_COUNT) return MADIPM_BAD_ARGUMENT;
    DeviceState* d = h->dev;
    if (vec == MADIPM_VEC_KKT_NZVAL) { d->image_valid = false; d->nzval_diag_stale = false; }
    LAUNCH(h, k_fill, dim3(vec_blocks(d->veclen[vec]), d->B), kVecThreads, 0, d->veclen[vec], d->vec[vec], d->veclen[vec], value);
    return MADIPM_OK;
}

int32_t madipm_vec_device_ptr(madipm_handle* h, int32_t vec, void** dev) {
    int32_t rc = check(h);
    if (rc) return rc;
    if (vec < 0 || vec >= MADIPM_VEC_COUNT || !dev) return MADIPM_BAD_ARGUMENT;
    if (vec == MADIPM_VEC_KKT_NZVAL) {                          // may be read or written through the pointer
        if ((rc = refresh_nzval_diag(h))) return rc;
        h->dev->image_valid = false;
    }
    *dev = h->dev->vec[vec];
    return MADIPM_OK;
}

int32_t madipm_synchronize(madipm_handle* h) {
    int32_t rc = check(h);
    if (rc) return rc;
    if ((rc = pending_enqueue(h))) return rc;
    { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }
    pending_collect(h);
    return MADIPM_OK;
}

int32_t madipm_debug_barrier_state(madipm_handle* h, int64_t* out8) {
    if (!h || !out8) return MADIPM_BAD_ARGUMENT;
    for (int k = 0; k < 8; ++k) out8[k] = 0;
    if (!h->dev || !h->dev->tree_dbg) return MADIPM_OK;
    const TreeDebug g = *h->dev->tree_dbg;
    const int64_t v[8] = {g.aborted, g.kernel, g.op, g.phase, g.block, g.grid, g.seen, g.target};
    for (int k = 0; k < 8; ++k) out8[k] = v[k];
    return MADIPM_OK;
}

const char* madipm_debug_describe(madipm_handle* h) {
    if (!h) return "";
    char buf[256];
    const unsigned pg = (h->dev && h->dev->tree_dbg) ? h->dev->tree_dbg->progress : 0u;
    std::snprintf(buf, sizeof(buf), "last kernel launched: %s (launch %lld); last tree progress mark: kernel %u op %u height %u; wedged %d; ",
                  h->last_launch, static_cast<long long>(h->launches), pg >> 24, (pg >> 16) & 0xffu, pg & 0xffffu, (h->dev && h->dev->wedged) ? 1 : 0);
    h->describe = std::string(buf) + (h->dev ? describe_tree_debug(h->dev) : std::string("no device state"));
    return h->describe.c_str();
}

int32_t madipm_debug_inject_barrier_fault(madipm_handle* h) {
    int32_t rc = check(h);
    if (rc) return rc;
    h->dev->inject_barrier_fault = true;
    return MADIPM_OK;
}

int32_t madipm_pinned_alloc(int64_t bytes, void** ptr) {
    if (!ptr || bytes < 0) return MADIPM_BAD_ARGUMENT;
    *ptr = nullptr;
    return cudaHostAlloc(ptr, static_cast<size_t>(std::max<int64_t>(bytes, 8)), cudaHostAllocPortable) == cudaSuccess ? MADIPM_OK : MADIPM_CUDA_ERROR;
}

int32_t madipm_pinned_free(void* ptr) {
    if (!ptr) return MADIPM_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? MADIPM_OK : MADIPM_CUDA_ERROR;
}

void* madipm_stream(madipm_handle* h) { return (h && h->dev) ? h->dev->stream : nullptr; }

int32_t madipm_vec_copy(madipm_handle* h, int32_t dst, int32_t src) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (dst < 0 || src < 0 || dst >= MADIPM_VEC_COUNT || src >= MADIPM_VEC_COUNT || d->veclen[dst] != d->veclen[src]) return MADIPM_BAD_ARGUMENT;
    if (src == MADIPM_VEC_KKT_NZVAL && (rc = refresh_nzval_diag(h))) return rc;
    if (dst == MADIPM_VEC_KKT_NZVAL) { d->image_valid = false; d->nzval_diag_stale = false; }
    CUDA_OK(h, cudaMemcpyAsync(d->vec[dst], d->vec[src], sizeof(double) * d->B * d->veclen[dst], cudaMemcpyDeviceToDevice, d->stream));
    return MADIPM_OK;
}

int32_t madipm_vec_axpy(madipm_handle* h, int32_t dst, int64_t off_dst, int32_t src, int64_t off_src, int64_t len, double alpha) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (dst < 0 || src < 0 || dst >= MADIPM_VEC_COUNT || src >= MADIPM_VEC_COUNT || len < 0 || off_dst < 0 || off_src < 0 ||
        off_dst + len > d->veclen[dst] || off_src + len > d->veclen[src]) return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_VECTOR);
    LAUNCH(h, k_axpy, dim3(vec_blocks(len), d->B), kVecThreads, 0, len, d->vec[dst] + off_dst, d->veclen[dst], d->vec[src] + off_src, d->veclen[src], alpha);
    return MADIPM_OK;
}

int32_t madipm_vec_norm_inf(madipm_handle* h, int32_t vec, double* out) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (vec < 0 || vec >= MADIPM_VEC_COUNT || !out) return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_VECTOR);
    LAUNCH(h, k_norm_inf, dim3(vec_blocks(d->veclen[vec]), d->B), kVecThreads, 0, d->veclen[vec], d->vec[vec], d->veclen[vec], d->red_scratch, d->red_counter, d->d_out);
    return fetch(h, d->d_out, out, d->B);
}

// ---- sub-range and index-view operations (the vector type of julia/MadIPMB200.jl) ---------------------------
static bool range_ok(const DeviceState* d, int32_t vec, int64_t off, int64_t len) {
    return vec >= 0 && vec < MADIPM_VEC_COUNT && off >= 0 && len >= 0 && off + len <= d->veclen[vec];
}

int32_t madipm_vec_get_range(madipm_handle* h, int32_t vec, int64_t off, int64_t len, double* host) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!host || !range_ok(d, vec, off, len)) return MADIPM_BAD_ARGUMENT;
    if (vec == MADIPM_VEC_KKT_NZVAL && (rc = refresh_nzval_diag(h))) return rc;
    if ((rc = sync_stream(h))) return rc;
    if (len > 0)
        CUDA_OK(h, cudaMemcpy2DAsync(host, sizeof(double) * len, d->vec[vec] + off, sizeof(double) * d->veclen[vec], sizeof(double) * len,
                                     d->B, cudaMemcpyDeviceToHost, d->stream));
    return sync_stream(h);
}

int32_t madipm_vec_set_range(madipm_handle* h, int32_t vec, int64_t off, int64_t len, const double* host) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!host || !range_ok(d, vec, off, len)) return MADIPM_BAD_ARGUMENT;
    if (vec == MADIPM_VEC_KKT_NZVAL) { d->image_valid = false; d->nzval_diag_stale = false; }
    if ((rc = sync_stream(h))) return rc;
    if (len > 0)
        CUDA_OK(h, cudaMemcpy2DAsync(d->vec[vec] + off, sizeof(double) * d->veclen[vec], host, sizeof(double) * len, sizeof(double) * len,
                                     d->B, cudaMemcpyHostToDevice, d->stream));
    return sync_stream(h);                                  // the host array may be reused as soon as the call returns
}

int32_t madipm_vec_fill_range(madipm_handle* h, int32_t vec, int64_t off, int64_t len, double value) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!range_ok(d, vec, off, len)) return MADIPM_BAD_ARGUMENT;
    if (vec == MADIPM_VEC_KKT_NZVAL) { d->image_valid = false; d->nzval_diag_stale = false; }
    if (len > 0) LAUNCH(h, k_fill, dim3(vec_blocks(len), d->B), kVecThreads, 0, len, d->vec[vec] + off, d->veclen[vec], value);
    return MADIPM_OK;
}

int32_t madipm_vec_copy_range(madipm_handle* h, int32_t dst, int64_t off_dst, int32_t src, int64_t off_src, int64_t len) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!range_ok(d, dst, off_dst, len) || !range_ok(d, src, off_src, len)) return MADIPM_BAD_ARGUMENT;
    if (src == MADIPM_VEC_KKT_NZVAL && (rc = refresh_nzval_diag(h))) return rc;
    if (dst == MADIPM_VEC_KKT_NZVAL) { d->image_valid = false; d->nzval_diag_stale = false; }
    if (len > 0 && !(dst == src && off_dst == off_src))
        LAUNCH(h, k_copy, dim3(vec_blocks(len), d->B), kVecThreads, 0, len, d->vec[dst] + off_dst, d->veclen[dst], d->vec[src] + off_src,
               d->veclen[src]);
    return MADIPM_OK;
}

int32_t madipm_vec_norm_inf_range(madipm_handle* h, int32_t vec, int64_t off, int64_t len, double* out) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!out || !range_ok(d, vec, off, len)) return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_VECTOR);
    LAUNCH(h, k_norm_inf, dim3(vec_blocks(len), d->B), kVecThreads, 0, len, d->vec[vec] + off, d->veclen[vec], d->red_scratch, d->red_counter, d->d_out);
    return fetch(h, d->d_out, out, d->B);
}

int32_t madipm_vec_axpy_at(madipm_handle* h, int32_t dst, int32_t which, int32_t src, int64_t off_src, double alpha) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (which < 0 || which > 1 || dst < 0 || dst >= MADIPM_VEC_COUNT || d->veclen[dst] != d->n) return MADIPM_BAD_ARGUMENT;
    const int cnt = which == 0 ? d->nlb : d->nub;
    if (!range_ok(d, src, off_src, cnt)) return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_VECTOR);
    if (cnt > 0)
        LAUNCH(h, k_axpy_at, dim3(vec_blocks(cnt), d->B), kVecThreads, 0, cnt, which == 0 ? d->ind_lb : d->ind_ub, d->vec[dst], (int64_t)d->n,
               d->vec[src] + off_src, d->veclen[src], alpha);
    return MADIPM_OK;
}

int32_t madipm_adjust_boundary(madipm_handle* h, const double* mu) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!mu) return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_VECTOR);
    CUDA_OK(h, cudaMemcpyAsync(d->d_mu, mu, sizeof(double) * d->B, cudaMemcpyHostToDevice, d->stream));
    LAUNCH(h, k_adjust_boundary, dim3(vec_blocks(d->n), d->B), kVecThreads, 0, make_view(*d), d->d_mu);
    return MADIPM_OK;
}

// ---- KKT system ---------------------------------------------------------------------------------
int32_t madipm_kkt_initialize(madipm_handle* h) {
    int32_t rc = check(h);
    if (rc) return rc;
    const struct { int v; double x; } init[] = {
        {MADIPM_VEC_REG, 1.0}, {MADIPM_VEC_PR_DIAG, 1.0}, {MADIPM_VEC_DU_DIAG, 0.0}, {MADIPM_VEC_L_LOWER, 0.0},
        {MADIPM_VEC_U_LOWER, 0.0}, {MADIPM_VEC_L_DIAG, 1.0}, {MADIPM_VEC_U_DIAG, 1.0}, {MADIPM_VEC_SCALING, 1.0}};
    for (auto& it : init)
        if ((rc = madipm_vec_fill(h, it.v, it.x))) return rc;
    return MADIPM_OK;
}

int32_t madipm_set_aug_diagonal_reg(madipm_handle* h, double del_w, double del_c) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    Timer t(h, MADIPM_T_ASSEMBLY);
    if (d->scaled)
        LAUNCH(h, k_set_aug_diagonal_scaled, dim3(vec_blocks(std::max(d->n, d->m)), d->B), kVecThreads, 0, make_view(*d), del_w, del_c);
    else
        LAUNCH(h, k_set_aug_diagonal_reg, dim3(vec_blocks(std::max(d->n, d->m)), d->B), kVecThreads, 0, make_view(*d), del_w, del_c);
    return MADIPM_OK;
}

int32_t madipm_set_initial_regularization(madipm_handle* h, double del_w, double del_c) {
    int32_t rc;
    if ((rc = madipm_vec_fill(h, MADIPM_VEC_REG, del_w))) return rc;
    if ((rc = madipm_vec_fill(h, MADIPM_VEC_PR_DIAG, del_w))) return rc;
    return madipm_vec_fill(h, MADIPM_VEC_DU_DIAG, del_c);
}

int32_t madipm_build_kkt(madipm_handle* h) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    Timer t(h, MADIPM_T_ASSEMBLY);
    double* nz = d->vec[MADIPM_VEC_KKT_NZVAL];
    if (d->scaled) {
        LAUNCH(h, k_build_kkt_k25, dim3(vec_blocks(d->kkt_nnz), d->B), kVecThreads, 0, d->kkt_nnz, d->n, d->m, d->kkt_row, d->kkt_col,
               d->nz_base, d->vec[MADIPM_VEC_SCALING], d->vec[MADIPM_VEC_PR_DIAG], d->vec[MADIPM_VEC_DU_DIAG], nz);
        d->nzval_diag_stale = false;
    } else if (d->kind == 0 && d->image_valid) {
        LAUNCH(h, k_build_kkt_k2_perm, dim3(vec_blocks(d->N), d->B), kVecThreads, 0, d->n, d->m, d->perm,
               d->vec[MADIPM_VEC_PR_DIAG], d->vec[MADIPM_VEC_DU_DIAG], d->dperm);
        d->nzval_diag_stale = true;
    } else if (d->kind == 0) {
        LAUNCH(h, k_build_kkt_k2, dim3(vec_blocks(d->N), d->B), kVecThreads, 0, d->n, d->m, d->diag_slot, d->const_diag,
               d->vec[MADIPM_VEC_PR_DIAG], d->vec[MADIPM_VEC_DU_DIAG], nz, d->kkt_nnz);
        d->nzval_diag_stale = false;
    } else {
        LAUNCH(h, k_assemble_normal, dim3(vec_blocks(d->kkt_nnz), d->B), kVecThreads, 0, d->kkt_nnz, d->prod_ptr, d->prod_a,
               d->prod_b, d->prod_k, d->A_val, d->nnzA, d->vec[MADIPM_VEC_PR_DIAG], (int64_t)d->n, nz, d->kkt_nnz);
    }
    return MADIPM_OK;
}

// Everything madipm_factorize puts on the stream.  fuse_forward: the tree launches are the factor + forward variant
// (kernels_tree.cuh OP 3): the right-hand side waiting in xw gets its forward sweep on the way.  Only when every supernode
// is small (no multi-block fronts); the caller checks.
static int32_t factorize_enqueue(madipm_handle* h, bool fuse_forward, bool* timed_tree) {
    int32_t rc;
    DeviceState* d = h->dev;
    const int B = d->B;
    const bool light = d->profiling == 2;
    bool first_tree = true;
    if (light) cudaEventRecord(d->ev0, d->stream);
    CUDA_OK(h, cudaMemsetAsync(d->fail, 0, sizeof(int) * B, d->stream));
    if (!d->image_valid) {                             // (with a valid panel image the tasks read P0 instead)
        CUDA_OK(h, cudaMemsetAsync(d->Lx, 0, sizeof(double) * B * d->lx_stride, d->stream));
        LAUNCH(h, k_assemble_panels, dim3(vec_blocks(d->kkt_nnz) * 4, B), kVecThreads, 0, d->kkt_nnz, d->asm_map,
               d->vec[MADIPM_VEC_KKT_NZVAL], d->kkt_nnz, d->Lx, d->lx_stride);
    }
    for (const DeviceState::TreeSeg& sg : d->segs) {
        if (light && first_tree) cudaEventRecord(d->ev2, d->stream);
        if ((rc = fuse_forward ? launch_tree<3>(h, sg) : launch_tree<0>(h, sg))) return rc;
        if (light && first_tree) { cudaEventRecord(d->ev3, d->stream); first_tree = false; }
        const int l = sg.ph1 - 1;
        const int nbig = d->big_ptr[l + 1] - d->big_ptr[l];
        if (nbig > 0) {
            const FactorView F = make_factor_view(*d, 0);
            const int* list = d->big_sn + d->big_ptr[l];
            const int exclusive = nbig == 1 ? 1 : 0;     // the only writer of its targets during these launches
            if (d->big_use_panel) {
                LAUNCH(h, k_big_panel, dim3(std::max(1, d->big_panelblocks[l]), nbig, B), kPanelThreads, d->smem_big_panel, F, list,
                       d->big_zbuf, d->big_zstride);
                LAUNCH(h, k_big_store_diag, dim3(nbig, B), 256, 0, F, list, d->big_zbuf, d->big_zstride);
            } else {
                LAUNCH(h, k_big_diag, dim3(nbig, B), 256, d->smem_big_diag, F, list, d->big_zbuf, d->big_zstride);
                if (d->big_rowblocks[l] > 0)
                    LAUNCH(h, k_big_trsm, dim3(d->big_rowblocks[l], nbig, B), 128, d->smem_big_trsm, F, list, d->big_zbuf, d->big_zstride);
            }
            if (d->big_rowblocks[l] > 0) {
                const int b0 = d->big_ptr[l];
                LAUNCH(h, k_big_update, dim3(d->big_tiles[l], nbig, B), 128, d->smem_big_update, F, list, d->upd_lim + b0,
                       d->upd_nsrc + b0, d->upd_src + static_cast<size_t>(b0) * kBigGroup, exclusive);
            }
        }
    }
    if (light) cudaEventRecord(d->ev1, d->stream);
    *timed_tree = !first_tree;
    return MADIPM_OK;
}

int32_t madipm_factorize(madipm_handle* h, int32_t* n_failed) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    const int B = d->B;
    const bool light = d->profiling == 2;
    bool timed_tree = false;
    {
        Timer t(h, MADIPM_T_FACTOR);
        if ((rc = factorize_enqueue(h, false, &timed_tree))) return rc;
        if (!n_failed) {                                   // asynchronous: flags and event times ride on the next synchronisation
            d->fail_pending = true;
            d->prof_pending = light;
            d->prof_tree_pending = light && timed_tree;
            return MADIPM_OK;
        }
        d->fail_pending = d->prof_pending = d->prof_tree_pending = false;
        { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }      // (pageable destination: guard first)
        CUDA_OK(h, cudaMemcpyAsync(d->h_fail.data(), d->fail, sizeof(int) * B, cudaMemcpyDeviceToHost, d->stream));
        { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }
        if (light) {
            float ms = 0;
            cudaEventElapsedTime(&ms, d->ev0, d->ev1);
            d->t_ms[MADIPM_T_FACTOR] += ms; d->t_calls[MADIPM_T_FACTOR] += 1;
            if (timed_tree) {
                cudaEventElapsedTime(&ms, d->ev2, d->ev3);
                d->t_ms[MADIPM_T_FACTOR_TREE] += ms; d->t_calls[MADIPM_T_FACTOR_TREE] += 1;
                if (cudaEventElapsedTime(&ms, d->ev4, d->ev3) == cudaSuccess) { d->t_ms[MADIPM_T_FACTOR_TOP] += ms; d->t_calls[MADIPM_T_FACTOR_TOP] += 1; }
            }
        }
    }
    CUDA_OK(h, cudaGetLastError());
    int bad = 0;
    for (int i = 0; i < B; ++i) bad += d->h_fail[i] != 0;
    if (n_failed) *n_failed = bad;
    return bad ? MADIPM_FACTOR_BREAKDOWN : MADIPM_OK;
}

int32_t madipm_set_pivot_checks(madipm_handle* h, double tiny, double growth) {
    int32_t rc = check(h);
    if (rc) return rc;
    if (!(tiny >= 0.0) || !(growth >= 0.0)) return MADIPM_BAD_ARGUMENT;
    h->dev->tiny_pivot = tiny;
    h->dev->growth_limit = growth;
    return MADIPM_OK;
}

int32_t madipm_set_tree_dataflow(madipm_handle* h, int32_t on, int32_t* in_effect) {
    int32_t rc = check(h);
    if (rc) return rc;
    h->dev->dataflow = on != 0;
    if (in_effect) *in_effect = (h->dev->dataflow && h->dev->dataflow_ok) ? 1 : 0;
    return MADIPM_OK;
}

int32_t madipm_is_factorized(madipm_handle* h, int32_t* flags) {
    int32_t rc = check(h);
    if (rc) return rc;
    if (!flags) return MADIPM_BAD_ARGUMENT;
    if (h->dev->fail_pending || h->dev->prof_pending) {
        if ((rc = pending_enqueue(h))) return rc;
        { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }
        pending_collect(h);
    }
    for (int i = 0; i < h->dev->B; ++i) flags[i] = h->dev->h_fail[i] == 0;
    return MADIPM_OK;
}

int32_t madipm_inertia(madipm_handle* h, int32_t inst, int64_t* n_pos, int64_t* n_zero, int64_t* n_neg) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (inst < 0 || inst >= d->B) return MADIPM_BAD_ARGUMENT;
    CUDA_OK(h, cudaMemsetAsync(d->d_inertia, 0, sizeof(unsigned long long) * 3 * d->B, d->stream));
    LAUNCH(h, k_inertia, dim3(vec_blocks(d->N), d->B), kVecThreads, 0, d->N, d->diag_pos, d->Lx, d->lx_stride, d->d_inertia);
    std::vector<unsigned long long> out(3 * d->B);
    { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }
    CUDA_OK(h, cudaMemcpyAsync(out.data(), d->d_inertia, sizeof(unsigned long long) * 3 * d->B, cudaMemcpyDeviceToHost, d->stream));
    { int32_t rc_ = sync_stream(h); if (rc_) return rc_; }
    if (n_pos) *n_pos = out[inst * 3];
    if (n_zero) *n_zero = out[inst * 3 + 1];
    if (n_neg) *n_neg = out[inst * 3 + 2];
    return MADIPM_OK;
}

int32_t madipm_linear_solve(madipm_handle* h, double* dev_rhs, int64_t stride) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!dev_rhs) return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_SOLVE);
    LAUNCH(h, k_perm_gather, dim3(vec_blocks(d->N), d->B), kVecThreads, 0, d->N, d->perm, dev_rhs, stride, d->xw, (int64_t)d->N);
    if ((rc = tri_solve(h))) return rc;
    LAUNCH(h, k_perm_scatter, dim3(vec_blocks(d->N), d->B), kVecThreads, 0, d->N, d->perm, d->xw, (int64_t)d->N, dev_rhs, stride);
    return MADIPM_OK;
}

// K4 / K4' in stages: 1 = everything before the triangular sweeps (reduce_rhs!, elimination, permutation into xw),
// 2 = backward sweep + everything after it, 3 = all of it.  Stage 1 needs nothing from the factorisation, which is what lets
// the predictor's forward sweep ride inside the factor launch (madipm_phase_factor_predictor).
static int32_t kkt_solve_stages(madipm_handle* h, int32_t vec_w, int stages) {
    int32_t rc;
    DeviceState* d = h->dev;
    if (vec_w < MADIPM_VEC_D || vec_w > MADIPM_VEC_W2) return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_SOLVE);
    VecView V = make_view(*d);
    double* w = d->vec[vec_w];
    const dim3 gN(vec_blocks(d->N), d->B), gn(vec_blocks(d->n), d->B), gm(vec_blocks(d->m), d->B);
    const bool pre = stages & 1, post = stages & 2;
    if (d->scaled) {
        if (pre) LAUNCH(h, k_k25_presolve, gN, kVecThreads, 0, V, w, d->perm, d->N, d->xw, (int64_t)d->N);
        if (post) {
            if ((rc = tri_solve(h, pre))) return rc;
            LAUNCH(h, k_k25_postsolve, gN, kVecThreads, 0, V, w, d->iperm, d->N, d->xw, (int64_t)d->N);
        }
    } else if (d->kind == 0) {
        if (pre) LAUNCH(h, k_k2_presolve, gN, kVecThreads, 0, V, w, d->perm, d->N, d->xw, (int64_t)d->N);
        if (post) {
            if ((rc = tri_solve(h, pre))) return rc;
            LAUNCH(h, k_k2_postsolve, gN, kVecThreads, 0, V, w, d->iperm, d->N, d->xw, (int64_t)d->N);
        }
    } else {
        if (pre) {
            LAUNCH(h, k_normal_pre, gn, kVecThreads, 0, V, w, d->buffer_n, (int64_t)d->n);
            LAUNCH(h, k_normal_rhs, gm, kVecThreads, 0, V, d->A_rp, d->A_ci, d->A_val, d->nnzA, w, d->buffer_n, (int64_t)d->n, d->perm, d->xw, (int64_t)d->N);
        }
        if (post) {
            if ((rc = tri_solve(h, pre))) return rc;
            LAUNCH(h, k_normal_unperm, gm, kVecThreads, 0, V, w, d->iperm, d->xw, (int64_t)d->N);
            LAUNCH(h, k_normal_post, gn, kVecThreads, 0, V, d->AT_rp, d->AT_ci, d->AT_val, d->nnzA, w);
        }
    }
    return MADIPM_OK;
}

int32_t madipm_kkt_solve(madipm_handle* h, int32_t vec_w) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    return kkt_solve_stages(h, vec_w, 3);
}

int32_t madipm_kkt_mul(madipm_handle* h, int32_t vec_w, int32_t vec_v, double alpha, double beta) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (vec_w < MADIPM_VEC_D || vec_w > MADIPM_VEC_W2 || vec_v < MADIPM_VEC_D || vec_v > MADIPM_VEC_W2 || vec_w == vec_v) return MADIPM_BAD_ARGUMENT;
    Timer t(h, MADIPM_T_SPMV);
    double* w = d->vec[vec_w];
    const double* v = d->vec[vec_v];
    // wx = alpha * (H vx + A' vy) + beta * wx ;  wy = alpha * A vx + beta * wy
    if (d->kind == 0 && d->nnzQ > 0) {
        spmv_launch(h, d->n, d->nnzQ, d->Q_rp, d->Q_ci, d->Q_val, d->nnzQ, v, d->L, w, d->L, alpha, beta);
        spmv_launch(h, d->n, d->nnzA, d->AT_rp, d->AT_ci, d->AT_val, d->nnzA, v + d->n, d->L, w, d->L, alpha, 1.0);
    } else {
        spmv_launch(h, d->n, d->nnzA, d->AT_rp, d->AT_ci, d->AT_val, d->nnzA, v + d->n, d->L, w, d->L, alpha, beta);
    }
    spmv_launch(h, d->m, d->nnzA, d->A_rp, d->A_ci, d->A_val, d->nnzA, v, d->L, w + d->n, d->L, alpha, beta);
    LAUNCH(h, k_kktmul_tail, dim3(vec_blocks(std::max(d->n, d->m)), d->B), kVecThreads, 0, make_view(*d), w, v, alpha, beta);
    return MADIPM_OK;
}

// solve_system! up to the three norms, which stay on the device in `norms_dev` (3 per instance)
// forward_done: d <- p and the solve's stage 1 were enqueued earlier and the forward sweep rode in the factor launch
static int32_t solve_system_enqueue(madipm_handle* h, int32_t vec_d, int32_t vec_p, double* norms_dev, bool forward_done = false) {
    int32_t rc;
    DeviceState* d = h->dev;
    if (!forward_done) {
        if ((rc = madipm_vec_copy(h, vec_d, vec_p))) return rc;             // linear_solver.jl:25
        if ((rc = madipm_kkt_solve(h, vec_d))) return rc;                   // :26
    } else if ((rc = kkt_solve_stages(h, vec_d, 2))) return rc;
    if ((rc = madipm_vec_copy(h, MADIPM_VEC_W1, vec_p))) return rc;         // :30
    if ((rc = madipm_kkt_mul(h, MADIPM_VEC_W1, vec_d, -1.0, 1.0))) return rc;   // :31
    Timer t(h, MADIPM_T_VECTOR);
    LAUNCH(h, k_norms3, dim3(vec_blocks(d->L), d->B), kVecThreads, 0, d->L, d->vec[MADIPM_VEC_W1], d->vec[vec_p],
           d->vec[vec_d], d->L, d->red_scratch, d->red_counter, norms_dev);
    return MADIPM_OK;
}

static int32_t norms_status(const DeviceState* d, const double* norms) {
    for (int i = 0; i < d->B; ++i) {
        double ratio = norms[3 * i] / std::max(1.0, norms[3 * i + 1]);
        if (std::isnan(ratio)) return MADIPM_NONFINITE;                     // :41-43
    }
    return MADIPM_OK;
}

int32_t madipm_solve_system(madipm_handle* h, int32_t vec_d, int32_t vec_p, double* norms) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!norms || vec_d == MADIPM_VEC_W1 || vec_p == MADIPM_VEC_W1) return MADIPM_BAD_ARGUMENT;
    if ((rc = solve_system_enqueue(h, vec_d, vec_p, d->d_out))) return rc;
    if ((rc = fetch(h, d->d_out, norms, 3 * d->B))) return rc;
    return norms_status(d, norms);
}

int32_t madipm_spmv_ptr(madipm_handle* h, int32_t op, double* dev_out, int64_t so, const double* dev_in, int64_t si) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    Timer t(h, MADIPM_T_SPMV);
    switch (op) {
        case MADIPM_OP_A: return spmv_launch(h, d->m, d->nnzA, d->A_rp, d->A_ci, d->A_val, d->nnzA, dev_in, si, dev_out, so, 1.0, 0.0);
        case MADIPM_OP_AT: return spmv_launch(h, d->n, d->nnzA, d->AT_rp, d->AT_ci, d->AT_val, d->nnzA, dev_in, si, dev_out, so, 1.0, 0.0);
        case MADIPM_OP_Q: return spmv_launch(h, d->n, d->nnzQ, d->Q_rp, d->Q_ci, d->Q_val, d->nnzQ, dev_in, si, dev_out, so, 1.0, 0.0);
        default: return MADIPM_BAD_ARGUMENT;
    }
}

int32_t madipm_spmv(madipm_handle* h, int32_t op, int32_t vec_out, int32_t vec_in) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (vec_out < 0 || vec_in < 0 || vec_out >= MADIPM_VEC_COUNT || vec_in >= MADIPM_VEC_COUNT || vec_out == vec_in) return MADIPM_BAD_ARGUMENT;
    const int64_t want_in = (op == MADIPM_OP_AT) ? d->m : d->n, want_out = (op == MADIPM_OP_A) ? d->m : d->n;
    if (d->veclen[vec_in] < want_in || d->veclen[vec_out] < want_out) return MADIPM_BAD_ARGUMENT;
    return madipm_spmv_ptr(h, op, d->vec[vec_out], d->veclen[vec_out], d->vec[vec_in], d->veclen[vec_in]);
}

// ---- fused phases -------------------------------------------------------------------------------
int32_t madipm_eval_model(madipm_handle* h, double c0, double* obj) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    {
        Timer t(h, MADIPM_T_SPMV);
        LAUNCH(h, k_eval_model, dim3(vec_blocks(std::max(d->n, d->m)), d->B), kVecThreads, 0, make_view(*d), d->Q_rp, d->Q_ci,
               d->Q_val, d->nnzQ, d->A_rp, d->A_ci, d->A_val, d->nnzA, d->red_scratch, d->red_counter, d->d_obj);
    }
    if (obj) {
        std::vector<double> tmp(2 * d->B);
        if ((rc = fetch(h, d->d_obj, tmp.data(), 2 * d->B))) return rc;
        for (int i = 0; i < d->B; ++i) obj[i] = c0 + tmp[2 * i] + 0.5 * tmp[2 * i + 1];
    }
    return MADIPM_OK;
}

int32_t madipm_phase_residuals(madipm_handle* h, double* out) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!out) return MADIPM_BAD_ARGUMENT;
    {
        Timer t(h, MADIPM_T_SPMV);
        LAUNCH(h, k_residuals, dim3(vec_blocks(std::max(d->n, d->m)), d->B), kVecThreads, 0, make_view(*d), d->AT_rp, d->AT_ci,
               d->AT_val, d->nnzA, d->vec[MADIPM_VEC_D], d->red_scratch, d->red_counter, d->d_out);
    }
    return fetch(h, d->d_out, out, 4 * d->B);
}

// phase_residuals + the objective of the last madipm_eval_model(h, c0, NULL): one synchronisation for both
int32_t madipm_phase_residuals_obj(madipm_handle* h, double c0, double* out, double* obj) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!out || !obj) return MADIPM_BAD_ARGUMENT;
    {
        Timer t(h, MADIPM_T_SPMV);
        LAUNCH(h, k_residuals, dim3(vec_blocks(std::max(d->n, d->m)), d->B), kVecThreads, 0, make_view(*d), d->AT_rp, d->AT_ci,
               d->AT_val, d->nnzA, d->vec[MADIPM_VEC_D], d->red_scratch, d->red_counter, d->d_out);
    }
    double* hobj = d->h_out + 24 * d->B;
    CUDA_OK(h, cudaMemcpyAsync(hobj, d->d_obj, sizeof(double) * 2 * d->B, cudaMemcpyDeviceToHost, d->stream));
    if ((rc = fetch(h, d->d_out, out, 4 * d->B))) return rc;
    for (int i = 0; i < d->B; ++i) obj[i] = c0 + hobj[2 * i] + 0.5 * hobj[2 * i + 1];
    return MADIPM_OK;
}

static int32_t set_rhs(madipm_handle* h, int mode, const double* mu);
static void launch_ratio(madipm_handle* h, const double* tau_dev, double* out_dev);

// affine_direction! (solver.jl:294-297) with one synchronisation: set_predictive_rhs!, solve_system!(d, solver, p),
// ratio tests at tau = 1, affine complementarity, get_correction!.  norms[3] and out[4] per instance as in
// madipm_solve_system / madipm_phase_affine.
static int32_t predictor_tail(madipm_handle* h, double* norms, double* out, bool forward_done) {
    int32_t rc;
    DeviceState* d = h->dev;
    if ((rc = solve_system_enqueue(h, MADIPM_VEC_D, MADIPM_VEC_P, d->d_norms, forward_done))) return rc;
    {
        Timer t(h, MADIPM_T_VECTOR);
        launch_ratio(h, d->d_ones, d->d_out2);                                                   // tau = 1 (solver.jl:295)
        LAUNCH(h, k_steps_from_ratio, (d->B + 127) / 128, 128, 0, d->B, d->d_out2, d->d_steps);
        LAUNCH(h, k_complementarity, dim3(vec_blocks(std::max(d->nlb, d->nub)), d->B), kVecThreads, 0, make_view(*d),
               d->vec[MADIPM_VEC_D], d->d_steps, 1, d->red_scratch, d->red_counter, d->d_out);   // solver.jl:296-297
    }
    double* hst = d->h_out + 16 * d->B;
    double* hno = d->h_out + 20 * d->B;
    std::vector<double> tmp(4 * d->B);
    CUDA_OK(h, cudaMemcpyAsync(hst, d->d_steps, sizeof(double) * 2 * d->B, cudaMemcpyDeviceToHost, d->stream));
    CUDA_OK(h, cudaMemcpyAsync(hno, d->d_norms, sizeof(double) * 3 * d->B, cudaMemcpyDeviceToHost, d->stream));
    if ((rc = fetch(h, d->d_out, tmp.data(), 4 * d->B))) return rc;
    std::memcpy(norms, hno, sizeof(double) * 3 * d->B);
    const int nb = d->nlb + d->nub;
    for (int i = 0; i < d->B; ++i) {
        out[4 * i] = hst[2 * i];
        out[4 * i + 1] = hst[2 * i + 1];
        out[4 * i + 2] = nb ? (tmp[4 * i] + tmp[4 * i + 1]) / nb : 0.0;
        out[4 * i + 3] = nb ? (tmp[4 * i + 2] + tmp[4 * i + 3]) / nb : 0.0;
    }
    return norms_status(d, norms);
}

int32_t madipm_phase_predictor(madipm_handle* h, double* norms, double* out) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    if (!norms || !out) return MADIPM_BAD_ARGUMENT;
    if ((rc = set_rhs(h, 2, nullptr))) return rc;
    return predictor_tail(h, norms, out, false);
}

// First try of factorize_regularized_system! (linear_solver.jl:6-17) + affine_direction! + the scalars of solver.jl:295-297
// with ONE synchronisation and the predictor's forward sweep inside the factor launch: set_aug_diagonal_reg!, build_kkt!,
// set_predictive_rhs!, d <- p, the pre-sweep part of solve! (none of which needs the factors), then factorisation with the
// forward step of every supernode done while its panel is still in shared memory, backward sweep, the rest of solve_system!.
// Same kernels' arithmetic as the separate calls (the forward step is the same device function): identical results.
// Breakdown flags as after madipm_factorize(h, NULL): read them with madipm_is_factorized; on a breakdown the caller
// re-enters the reference's retry (madipm_set_aug_diagonal_reg ... madipm_phase_predictor).
int32_t madipm_phase_factor_predictor(madipm_handle* h, double del_w, double del_c, double* norms, double* out) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!norms || !out) return MADIPM_BAD_ARGUMENT;
    if ((rc = madipm_set_aug_diagonal_reg(h, del_w, del_c))) return rc;
    if ((rc = madipm_build_kkt(h))) return rc;
    const bool fuse = d->sym->n_big == 0 && !d->segs.empty();
    if (fuse) {
        if ((rc = set_rhs(h, 2, nullptr))) return rc;
        if ((rc = madipm_vec_copy(h, MADIPM_VEC_D, MADIPM_VEC_P))) return rc;                    // linear_solver.jl:25
        if ((rc = kkt_solve_stages(h, MADIPM_VEC_D, 1))) return rc;
    }
    {
        Timer t(h, MADIPM_T_FACTOR);
        bool timed_tree = false;
        if ((rc = factorize_enqueue(h, fuse, &timed_tree))) return rc;
        d->fail_pending = true;
        d->prof_pending = d->profiling == 2;
        d->prof_tree_pending = d->prof_pending && timed_tree;
    }
    if (!fuse && (rc = set_rhs(h, 2, nullptr))) return rc;
    return predictor_tail(h, norms, out, fuse);
}

int32_t madipm_phase_iteration(madipm_handle* h, const madipm_iter_options* o, const int32_t* active, double* out, int32_t* ok) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!o || !out || !ok) return MADIPM_BAD_ARGUMENT;
    const int B = d->B;
    double* d_aff4 = d->d_iter;
    double* d_norms2 = d_aff4 + 4 * B;
    double* d_ratio8 = d_norms2 + 3 * B;
    double* d_res4 = d_ratio8 + 8 * B;
    const int* act = nullptr;
    if (active) {
        CUDA_OK(h, cudaMemcpyAsync(d->d_active, active, sizeof(int) * B, cudaMemcpyHostToDevice, d->stream));
        act = d->d_active;
    }
    // ---- factorize_regularized_system! (first try) + affine_direction!, as madipm_phase_factor_predictor
    if ((rc = madipm_set_aug_diagonal_reg(h, o->del_w, o->del_c))) return rc;
    if ((rc = madipm_build_kkt(h))) return rc;
    const bool fuse = d->sym->n_big == 0 && !d->segs.empty();
    if (fuse) {
        if ((rc = set_rhs(h, 2, nullptr))) return rc;
        if ((rc = madipm_vec_copy(h, MADIPM_VEC_D, MADIPM_VEC_P))) return rc;
        if ((rc = kkt_solve_stages(h, MADIPM_VEC_D, 1))) return rc;
    }
    {
        Timer t(h, MADIPM_T_FACTOR);
        bool timed_tree = false;
        if ((rc = factorize_enqueue(h, fuse, &timed_tree))) return rc;
        d->fail_pending = true;
        d->prof_pending = d->profiling == 2;
        d->prof_tree_pending = d->prof_pending && timed_tree;
    }
    if (!fuse && (rc = set_rhs(h, 2, nullptr))) return rc;
    if ((rc = solve_system_enqueue(h, MADIPM_VEC_D, MADIPM_VEC_P, d->d_norms, fuse))) return rc;
    const dim3 gb(vec_blocks(std::max(d->nlb, d->nub)), B);
    {
        Timer t(h, MADIPM_T_VECTOR);
        launch_ratio(h, d->d_ones, d->d_out2);                                                   // tau = 1 (solver.jl:295)
        LAUNCH(h, k_steps_from_ratio, (B + 127) / 128, 128, 0, B, d->d_out2, d->d_steps);
        LAUNCH(h, k_complementarity, gb, kVecThreads, 0, make_view(*d), d->vec[MADIPM_VEC_D], d->d_steps, 1, d->red_scratch,
               d->red_counter, d->d_out);                                                        // solver.jl:296-297
        // ---- update_barrier! and the step rule's tau, on the device
        LAUNCH(h, k_iter_scalars1, 1, 256, 0, B, d->nlb + d->nub, d->d_out, d->d_steps, d->d_norms, d->fail, act, o->has_inequalities,
               o->mu_min, o->adaptive_step, o->tau, o->res_tol, d_aff4, d->d_mu, d->d_tau, d->d_ok);
        // ---- mehrotra_correction_direction! (solver.jl:307-311)
        const int top = std::max(std::max(d->n, d->m), std::max(d->nlb, d->nub));
        LAUNCH(h, k_set_rhs, dim3(vec_blocks(top), B), kVecThreads, 0, make_view(*d), d->vec[MADIPM_VEC_P], 3, d->d_mu);
    }
    if ((rc = solve_system_enqueue(h, MADIPM_VEC_D, MADIPM_VEC_P, d_norms2))) return rc;
    {
        Timer t(h, MADIPM_T_VECTOR);
        launch_ratio(h, d->d_tau, d_ratio8);                                                     // update_step! (solver.jl:329)
        LAUNCH(h, k_iter_scalars2, 1, 256, 0, B, d_ratio8, d_norms2, act, d->d_steps, d->d_ok);
        // ---- solver.jl:332-335, 342: only if every active instance is fine
        LAUNCH(h, k_update_iterate, dim3(vec_blocks(std::max(d->n, d->m)), B), kVecThreads, 0, make_view(*d), d->vec[MADIPM_VEC_D],
               d->d_steps, d->d_mu, act, d->d_ok);
    }
    {
        Timer t(h, MADIPM_T_SPMV);                                                               // solver.jl:338-340, then :259-272
        LAUNCH(h, k_eval_model, dim3(vec_blocks(std::max(d->n, d->m)), B), kVecThreads, 0, make_view(*d), d->Q_rp, d->Q_ci,
               d->Q_val, d->nnzQ, d->A_rp, d->A_ci, d->A_val, d->nnzA, d->red_scratch, d->red_counter, d->d_obj);
        LAUNCH(h, k_residuals, dim3(vec_blocks(std::max(d->n, d->m)), B), kVecThreads, 0, make_view(*d), d->AT_rp, d->AT_ci,
               d->AT_val, d->nnzA, d->vec[MADIPM_VEC_D], d->red_scratch, d->red_counter, d_res4);
    }
    // ---- one synchronisation for everything
    double* hp = d->h_iter;
    double* h_norms1 = hp;              hp += 3 * B;
    double* h_aff4 = hp;                hp += 4 * B;
    double* h_mu = hp;                  hp += B;
    double* h_tau = hp;                 hp += B;
    double* h_norms2 = hp;              hp += 3 * B;
    double* h_ratio8 = hp;              hp += 8 * B;
    double* h_steps = hp;               hp += 2 * B;
    double* h_res4 = hp;                hp += 4 * B;
    double* h_obj2 = hp;                hp += 2 * B;
    int* h_ok = reinterpret_cast<int*>(hp);
    auto back = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d->stream); };
    CUDA_OK(h, back(h_norms1, d->d_norms, sizeof(double) * 3 * B));
    CUDA_OK(h, back(h_aff4, d_aff4, sizeof(double) * 4 * B));
    CUDA_OK(h, back(h_mu, d->d_mu, sizeof(double) * B));
    CUDA_OK(h, back(h_tau, d->d_tau, sizeof(double) * B));
    CUDA_OK(h, back(h_norms2, d_norms2, sizeof(double) * 3 * B));
    CUDA_OK(h, back(h_ratio8, d_ratio8, sizeof(double) * 8 * B));
    CUDA_OK(h, back(h_steps, d->d_steps, sizeof(double) * 2 * B));
    CUDA_OK(h, back(h_res4, d_res4, sizeof(double) * 4 * B));
    CUDA_OK(h, back(h_obj2, d->d_obj, sizeof(double) * 2 * B));
    CUDA_OK(h, back(h_ok, d->d_ok, sizeof(int)));
    if ((rc = pending_enqueue(h))) return rc;
    if ((rc = sync_stream(h))) return rc;
    CUDA_OK(h, cudaGetLastError());
    pending_collect(h);
    *ok = *h_ok;
    for (int i = 0; i < B; ++i) {
        double* r = out + 28 * i;
        for (int k = 0; k < 3; ++k) r[k] = h_norms1[3 * i + k];
        for (int k = 0; k < 4; ++k) r[3 + k] = h_aff4[4 * i + k];
        r[7] = h_mu[i];
        r[8] = h_tau[i];
        for (int k = 0; k < 3; ++k) r[9 + k] = h_norms2[3 * i + k];
        for (int k = 0; k < 8; ++k) r[12 + k] = h_ratio8[8 * i + k];
        r[20] = h_steps[2 * i];
        r[21] = h_steps[2 * i + 1];
        for (int k = 0; k < 4; ++k) r[22 + k] = h_res4[4 * i + k];
        r[26] = o->c0 + h_obj2[2 * i] + 0.5 * h_obj2[2 * i + 1];
        r[27] = 0.0;
    }
    return MADIPM_OK;
}

// mpc_direction! without Gondzio corrections (solver.jl:307-311) + the ratio tests of update_step! (solver.jl:329,
// kernels.jl:290-305) with one synchronisation: set_correction_rhs!(mu), solve_system!, ratio tests at tau.
int32_t madipm_phase_corrector(madipm_handle* h, const double* mu, const double* tau, double* norms, double* ratio) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!mu || !tau || !norms || !ratio) return MADIPM_BAD_ARGUMENT;
    if ((rc = set_rhs(h, 3, mu))) return rc;
    if ((rc = solve_system_enqueue(h, MADIPM_VEC_D, MADIPM_VEC_P, d->d_norms))) return rc;
    {
        Timer t(h, MADIPM_T_VECTOR);
        CUDA_OK(h, cudaMemcpyAsync(d->d_tau, tau, sizeof(double) * d->B, cudaMemcpyHostToDevice, d->stream));
        launch_ratio(h, d->d_tau, d->d_out);
    }
    double* hno = d->h_out + 20 * d->B;
    CUDA_OK(h, cudaMemcpyAsync(hno, d->d_norms, sizeof(double) * 3 * d->B, cudaMemcpyDeviceToHost, d->stream));
    if ((rc = fetch(h, d->d_out, ratio, 8 * d->B))) return rc;
    std::memcpy(norms, hno, sizeof(double) * 3 * d->B);
    return norms_status(d, norms);
}

static int32_t set_rhs(madipm_handle* h, int mode, const double* mu) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    Timer t(h, MADIPM_T_VECTOR);
    if (mode == 3) {
        if (!mu) return MADIPM_BAD_ARGUMENT;
        CUDA_OK(h, cudaMemcpyAsync(d->d_mu, mu, sizeof(double) * d->B, cudaMemcpyHostToDevice, d->stream));
    }
    const int top = std::max(std::max(d->n, d->m), std::max(d->nlb, d->nub));
    LAUNCH(h, k_set_rhs, dim3(vec_blocks(top), d->B), kVecThreads, 0, make_view(*d), d->vec[MADIPM_VEC_P], mode, d->d_mu);
    return MADIPM_OK;
}
int32_t madipm_set_initial_primal_rhs(madipm_handle* h) { return set_rhs(h, 0, nullptr); }
int32_t madipm_set_initial_dual_rhs(madipm_handle* h) { return set_rhs(h, 1, nullptr); }
int32_t madipm_set_predictive_rhs(madipm_handle* h) { return set_rhs(h, 2, nullptr); }
int32_t madipm_set_correction_rhs(madipm_handle* h, const double* mu) { return set_rhs(h, 3, mu); }

static void launch_ratio(madipm_handle* h, const double* tau_dev, double* out_dev) {
    DeviceState* d = h->dev;
    LAUNCH(h, k_ratio_test, dim3(vec_blocks(std::max(d->nlb, d->nub)), d->B), kVecThreads, 0, make_view(*d), d->vec[MADIPM_VEC_D],
           tau_dev, d->red_scratch, d->red_counter, out_dev);
}

int32_t madipm_ratio_test(madipm_handle* h, const double* tau, double* out) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!tau || !out) return MADIPM_BAD_ARGUMENT;
    {
        Timer t(h, MADIPM_T_VECTOR);
        CUDA_OK(h, cudaMemcpyAsync(d->d_tau, tau, sizeof(double) * d->B, cudaMemcpyHostToDevice, d->stream));
        launch_ratio(h, d->d_tau, d->d_out);
    }
    return fetch(h, d->d_out, out, 8 * d->B);
}

int32_t madipm_complementarity(madipm_handle* h, const double* alpha_p, const double* alpha_d, double* out) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!out) return MADIPM_BAD_ARGUMENT;
    std::vector<double> st(2 * d->B, 0.0);
    for (int i = 0; i < d->B; ++i) { st[2 * i] = alpha_p ? alpha_p[i] : 0.0; st[2 * i + 1] = alpha_d ? alpha_d[i] : 0.0; }
    {
        Timer t(h, MADIPM_T_VECTOR);
        CUDA_OK(h, cudaMemcpyAsync(d->d_steps, st.data(), sizeof(double) * 2 * d->B, cudaMemcpyHostToDevice, d->stream));
        LAUNCH(h, k_complementarity, dim3(vec_blocks(std::max(d->nlb, d->nub)), d->B), kVecThreads, 0, make_view(*d),
               d->vec[MADIPM_VEC_D], d->d_steps, 0, d->red_scratch, d->red_counter, d->d_out);
    }
    std::vector<double> tmp(4 * d->B);
    if ((rc = fetch(h, d->d_out, tmp.data(), 4 * d->B))) return rc;
    const int nb = d->nlb + d->nub;
    for (int i = 0; i < d->B; ++i) {
        out[2 * i] = nb ? (tmp[4 * i] + tmp[4 * i + 1]) / nb : 0.0;
        out[2 * i + 1] = nb ? (tmp[4 * i + 2] + tmp[4 * i + 3]) / nb : 0.0;
    }
    return MADIPM_OK;
}

int32_t madipm_phase_affine(madipm_handle* h, double* out) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!out) return MADIPM_BAD_ARGUMENT;
    {
        Timer t(h, MADIPM_T_VECTOR);
        launch_ratio(h, d->d_ones, d->d_out2);                                                   // tau = 1 (solver.jl:295)
        LAUNCH(h, k_steps_from_ratio, (d->B + 127) / 128, 128, 0, d->B, d->d_out2, d->d_steps);
        LAUNCH(h, k_complementarity, dim3(vec_blocks(std::max(d->nlb, d->nub)), d->B), kVecThreads, 0, make_view(*d),
               d->vec[MADIPM_VEC_D], d->d_steps, 1, d->red_scratch, d->red_counter, d->d_out);   // solver.jl:296-297
    }
    std::vector<double> st(2 * d->B), tmp(4 * d->B);
    CUDA_OK(h, cudaMemcpyAsync(d->h_out + 16 * d->B, d->d_steps, sizeof(double) * 2 * d->B, cudaMemcpyDeviceToHost, d->stream));
    if ((rc = fetch(h, d->d_out, tmp.data(), 4 * d->B))) return rc;
    std::memcpy(st.data(), d->h_out + 16 * d->B, sizeof(double) * 2 * d->B);
    const int nb = d->nlb + d->nub;
    for (int i = 0; i < d->B; ++i) {
        out[4 * i] = st[2 * i];
        out[4 * i + 1] = st[2 * i + 1];
        out[4 * i + 2] = nb ? (tmp[4 * i] + tmp[4 * i + 1]) / nb : 0.0;
        out[4 * i + 3] = nb ? (tmp[4 * i + 2] + tmp[4 * i + 3]) / nb : 0.0;
    }
    return MADIPM_OK;
}

int32_t madipm_get_correction(madipm_handle* h) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    Timer t(h, MADIPM_T_VECTOR);
    LAUNCH(h, k_complementarity, dim3(vec_blocks(std::max(d->nlb, d->nub)), d->B), kVecThreads, 0, make_view(*d),
           d->vec[MADIPM_VEC_D], (const double*)nullptr, 1, d->red_scratch, d->red_counter, d->d_out);
    return MADIPM_OK;
}

int32_t madipm_set_extra_correction(madipm_handle* h, const double* alpha_p, const double* alpha_d, double bmin, double bmax, const double* mu) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!alpha_p || !alpha_d || !mu) return MADIPM_BAD_ARGUMENT;
    std::vector<double> st(2 * d->B);
    for (int i = 0; i < d->B; ++i) { st[2 * i] = alpha_p[i]; st[2 * i + 1] = alpha_d[i]; }
    Timer t(h, MADIPM_T_VECTOR);
    CUDA_OK(h, cudaMemcpyAsync(d->d_steps, st.data(), sizeof(double) * 2 * d->B, cudaMemcpyHostToDevice, d->stream));
    CUDA_OK(h, cudaMemcpyAsync(d->d_mu, mu, sizeof(double) * d->B, cudaMemcpyHostToDevice, d->stream));
    LAUNCH(h, k_extra_correction, dim3(vec_blocks(std::max(d->nlb, d->nub)), d->B), kVecThreads, 0, make_view(*d),
           d->vec[MADIPM_VEC_D], d->d_steps, bmin, bmax, d->d_mu);
    return MADIPM_OK;
}

int32_t madipm_update_iterate(madipm_handle* h, const double* alpha_p, const double* alpha_d, const double* mu, const int32_t* active) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!alpha_p || !alpha_d || !mu) return MADIPM_BAD_ARGUMENT;
    std::vector<double> st(2 * d->B);
    for (int i = 0; i < d->B; ++i) { st[2 * i] = alpha_p[i]; st[2 * i + 1] = alpha_d[i]; }
    Timer t(h, MADIPM_T_VECTOR);
    CUDA_OK(h, cudaMemcpyAsync(d->d_steps, st.data(), sizeof(double) * 2 * d->B, cudaMemcpyHostToDevice, d->stream));
    CUDA_OK(h, cudaMemcpyAsync(d->d_mu, mu, sizeof(double) * d->B, cudaMemcpyHostToDevice, d->stream));
    if (active) CUDA_OK(h, cudaMemcpyAsync(d->d_active, active, sizeof(int) * d->B, cudaMemcpyHostToDevice, d->stream));
    LAUNCH(h, k_update_iterate, dim3(vec_blocks(std::max(d->n, d->m)), d->B), kVecThreads, 0, make_view(*d), d->vec[MADIPM_VEC_D],
           d->d_steps, d->d_mu, active ? d->d_active : (const int*)nullptr);
    return MADIPM_OK;
}

int32_t madipm_mehrotra_step(madipm_handle* h, double gamma_f, double* out) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!out) return MADIPM_BAD_ARGUMENT;
    {
        Timer t(h, MADIPM_T_VECTOR);
        launch_ratio(h, d->d_ones, d->d_out2);
        LAUNCH(h, k_steps_from_ratio, (d->B + 127) / 128, 128, 0, d->B, d->d_out2, d->d_steps);
        LAUNCH(h, k_complementarity, dim3(vec_blocks(std::max(d->nlb, d->nub)), d->B), kVecThreads, 0, make_view(*d),
               d->vec[MADIPM_VEC_D], d->d_steps, 0, d->red_scratch, d->red_counter, d->d_out);
        LAUNCH(h, k_mehrotra_step, (d->B + 127) / 128, 128, 0, make_view(*d), d->B, d->vec[MADIPM_VEC_D], d->d_out2, d->d_out,
               gamma_f, d->d_steps);
    }
    return fetch(h, d->d_steps, out, 2 * d->B);
}

int32_t madipm_init_multipliers_and_shift(madipm_handle* h, double bound_fac) {
    int32_t rc = check(h, true);
    if (rc) return rc;
    DeviceState* d = h->dev;
    const int B = d->B;
    VecView V = make_view(*d);
    const dim3 g(vec_blocks(d->n), B);
    // step 3 (solver.jl:35-39): res = A'y + f
    if ((rc = madipm_spmv(h, MADIPM_OP_AT, MADIPM_VEC_JACL, MADIPM_VEC_Y))) return rc;
    std::vector<double> r(8 * B), del(2 * B);
    LAUNCH(h, k_start_a, g, kVecThreads, 0, V, d->red_scratch, d->red_counter, d->d_out);
    if ((rc = fetch(h, d->d_out, r.data(), 4 * B))) return rc;
    for (int i = 0; i < B; ++i) {
        del[2 * i] = std::max(0.0, std::max(-1.5 * r[4 * i], -1.5 * r[4 * i + 1]));          // delta_x, l.68-72
        del[2 * i + 1] = std::max(0.0, std::max(-1.5 * r[4 * i + 2], -1.5 * r[4 * i + 3]));  // delta_s, l.74-78
    }
    CUDA_OK(h, cudaMemcpyAsync(d->d_steps, del.data(), sizeof(double) * 2 * B, cudaMemcpyHostToDevice, d->stream));
    LAUNCH(h, k_start_b, g, kVecThreads, 0, V, d->d_steps, d->red_scratch, d->red_counter, d->d_out);
    if ((rc = fetch(h, d->d_out, r.data(), 6 * B))) return rc;
    for (int i = 0; i < B; ++i) {
        const double* q = &r[6 * i];
        double mu = 0.0;
        if (d->nlb) mu += q[0];
        if (d->nub) mu += q[1];
        del[2 * i] = mu / (2 * (q[2] + q[3]));        // delta_x2, l.93
        del[2 * i + 1] = mu / (2 * (q[4] + q[5]));    // delta_s2, l.94
    }
    CUDA_OK(h, cudaMemcpyAsync(d->d_steps, del.data(), sizeof(double) * 2 * B, cudaMemcpyHostToDevice, d->stream));
    LAUNCH(h, k_start_c, g, kVecThreads, 0, V, d->d_steps, bound_fac, d->red_scratch, d->red_counter, d->d_out);
    if ((rc = fetch(h, d->d_out, r.data(), 4 * B))) return rc;
    for (int i = 0; i < B; ++i)
        for (int k = 0; k < 4; ++k)
            if (!(r[4 * i + k] > 0.0)) {   // the four @assert of solver.jl:120-123
                h->error = "starting point left the interior (assert of solver.jl:120-123)";
                return MADIPM_INTERNAL;
            }
    return MADIPM_OK;
}

// ---- device-resident inputs -----------------------------------------------------------------------
static const int kSnapVec[4] = {MADIPM_VEC_X, MADIPM_VEC_XL, MADIPM_VEC_XU, MADIPM_VEC_Y};

int32_t madipm_snapshot_inputs(madipm_handle* h) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    for (int k = 0; k < 4; ++k) {
        const size_t bytes = sizeof(double) * d->B * d->veclen[kSnapVec[k]];
        if (!d->snap[k] && (rc = dev_alloc(h, &d->snap[k], static_cast<size_t>(d->B) * d->veclen[kSnapVec[k]]))) return rc;
        CUDA_OK(h, cudaMemcpyAsync(d->snap[k], d->vec[kSnapVec[k]], bytes, cudaMemcpyDeviceToDevice, d->stream));
    }
    return MADIPM_OK;
}

int32_t madipm_restore_inputs(madipm_handle* h) {
    int32_t rc = check(h);
    if (rc) return rc;
    DeviceState* d = h->dev;
    if (!d->snap[0]) { h->error = "restore_inputs before snapshot_inputs"; return MADIPM_BAD_ARGUMENT; }
    for (int k = 0; k < 4; ++k)
        CUDA_OK(h, cudaMemcpyAsync(d->vec[kSnapVec[k]], d->snap[k], sizeof(double) * d->B * d->veclen[kSnapVec[k]],
                                   cudaMemcpyDeviceToDevice, d->stream));
    return MADIPM_OK;
}

// ---- measurement hooks ----------------------------------------------------------------------------
int32_t madipm_profile(madipm_handle* h, int32_t enable) {
    int32_t rc = check(h);
    if (rc) return rc;
    h->dev->profiling = enable;
    return MADIPM_OK;
}

int32_t madipm_profile_read(madipm_handle* h, double* ms, int64_t* calls, int32_t reset) {
    int32_t rc = check(h);
    if (rc) return rc;
    for (int i = 0; i < MADIPM_T_COUNT; ++i) {
        if (ms) ms[i] = h->dev->t_ms[i];
        if (calls) calls[i] = h->dev->t_calls[i];
        if (reset) { h->dev->t_ms[i] = 0; h->dev->t_calls[i] = 0; }
    }
    return MADIPM_OK;
}

}  // extern "C"
