// Host half of the C ABI: lifecycle, symbolic phase, getters.
// The numeric entry points live in device.cu.
#include <cstring>
#include <exception>

#include "handle.hpp"

using namespace madipm;

#define MADIPM_TRY(h, body)                                           \
    try { body; }                                                     \
    catch (const std::exception& e) { (h)->error = e.what(); return MADIPM_BAD_ARGUMENT; } \
    catch (...) { (h)->error = "unknown exception"; return MADIPM_INTERNAL; }

namespace {
template <class T>
int32_t copy_out(const std::vector<T>& v, int64_t* out, int64_t cap, int64_t* len) {
    if (len) *len = static_cast<int64_t>(v.size());
    if (!out) return MADIPM_OK;
    if (cap < static_cast<int64_t>(v.size())) return MADIPM_BAD_ARGUMENT;
    for (size_t i = 0; i < v.size(); ++i) out[i] = static_cast<int64_t>(v[i]);
    return MADIPM_OK;
}
}  // namespace

extern "C" {

const char* madipm_version(void) { return "madipm_b200 0.1.0 (sm_100a)"; }

void madipm_default_config(madipm_config* cfg) {
    if (!cfg) return;
    SymbolicOptions o;
    cfg->kkt_kind = MADIPM_KKT_K2;
    cfg->ordering = o.ordering;
    cfg->nd_leaf = o.nd_leaf;
    cfg->max_width = o.max_width;
    cfg->relax_small = o.relax_small;
    for (int i = 0; i < 3; ++i) cfg->relax_zeros[i] = o.relax_zeros[i];
    cfg->block_panel = o.block_panel;
    cfg->device = 0;
    cfg->batch = 1;
}

int32_t madipm_create(const madipm_config* cfg, madipm_handle** out) {
    if (!out) return MADIPM_BAD_ARGUMENT;
    madipm_handle* h = new (std::nothrow) madipm_handle();
    if (!h) return MADIPM_INTERNAL;
    if (cfg) h->cfg = *cfg;
    else madipm_default_config(&h->cfg);
    if (h->cfg.batch < 1 || h->cfg.max_width < 1 || h->cfg.max_width > 256 || h->cfg.kkt_kind < MADIPM_KKT_K2 ||
        h->cfg.kkt_kind > MADIPM_KKT_K2_SCALED) { delete h; return MADIPM_BAD_ARGUMENT; }
    *out = h;
    return MADIPM_OK;
}

int32_t madipm_destroy(madipm_handle* h) {
    if (!h) return MADIPM_OK;
    device_destroy(h);
    delete h;
    return MADIPM_OK;
}

const char* madipm_last_error(const madipm_handle* h) { return h ? h->error.c_str() : "null handle"; }

int32_t madipm_set_pattern_k2(madipm_handle* h, int32_t nvar, int32_t ncon,
                              int64_t nnzh, const int32_t* hess_I, const int32_t* hess_J,
                              int64_t nnzj, const int32_t* jac_I, const int32_t* jac_J,
                              int32_t n_slack, const int64_t* ind_ineq) {
    if (!h || nvar < 0 || ncon < 0 || nnzh < 0 || nnzj < 0 || n_slack < 0) return MADIPM_BAD_ARGUMENT;
    if (h->cfg.kkt_kind != MADIPM_KKT_K2 && h->cfg.kkt_kind != MADIPM_KKT_K2_SCALED) { h->error = "handle was configured for the normal-equations system"; return MADIPM_BAD_ARGUMENT; }
    MADIPM_TRY(h, build_k2_pattern(nvar, ncon, nnzh, hess_I, hess_J, nnzj, jac_I, jac_J, n_slack, ind_ineq, h->pat));
    h->has_pattern = true;
    h->analyzed = false;
    return MADIPM_OK;
}

int32_t madipm_set_pattern_normal(madipm_handle* h, int32_t nvar, int32_t ncon,
                                  int64_t nnzj, const int32_t* jac_I, const int32_t* jac_J,
                                  int32_t n_slack, const int64_t* ind_ineq) {
    if (!h || nvar < 0 || ncon < 0 || nnzj < 0 || n_slack < 0) return MADIPM_BAD_ARGUMENT;
    if (h->cfg.kkt_kind != MADIPM_KKT_NORMAL) { h->error = "handle was configured for the K2 system"; return MADIPM_BAD_ARGUMENT; }
    MADIPM_TRY(h, build_normal_kkt_pattern(nvar, ncon, nnzj, jac_I, jac_J, n_slack, ind_ineq, h->pat));
    h->has_pattern = true;
    h->analyzed = false;
    return MADIPM_OK;
}

int32_t madipm_set_bounds_index(madipm_handle* h, int32_t nlb, const int64_t* ind_lb, int32_t nub, const int64_t* ind_ub) {
    if (!h || nlb < 0 || nub < 0) return MADIPM_BAD_ARGUMENT;
    if (!h->has_pattern) { h->error = "set_bounds_index before set_pattern"; return MADIPM_BAD_ARGUMENT; }
    h->ind_lb.resize(nlb);
    h->ind_ub.resize(nub);
    for (int32_t i = 0; i < nlb; ++i) {
        if (ind_lb[i] < 1 || ind_lb[i] > h->pat.n_tot) { h->error = "ind_lb out of range"; return MADIPM_BAD_ARGUMENT; }
        h->ind_lb[i] = static_cast<int32_t>(ind_lb[i] - 1);
    }
    for (int32_t i = 0; i < nub; ++i) {
        if (ind_ub[i] < 1 || ind_ub[i] > h->pat.n_tot) { h->error = "ind_ub out of range"; return MADIPM_BAD_ARGUMENT; }
        h->ind_ub[i] = static_cast<int32_t>(ind_ub[i] - 1);
    }
    h->nlb = nlb;
    h->nub = nub;
    return MADIPM_OK;
}

int32_t madipm_analyze(madipm_handle* h) {
    if (!h) return MADIPM_BAD_ARGUMENT;
    if (!h->has_pattern) { h->error = "analyze before set_pattern"; return MADIPM_BAD_ARGUMENT; }
    SymbolicOptions o;
    o.ordering = h->cfg.ordering;
    o.nd_leaf = h->cfg.nd_leaf;
    o.max_width = h->cfg.max_width;
    o.relax_small = h->cfg.relax_small;
    for (int i = 0; i < 3; ++i) o.relax_zeros[i] = h->cfg.relax_zeros[i];
    if (h->cfg.block_panel > 0) o.block_panel = h->cfg.block_panel;
    std::vector<uint8_t> delayed;
    if (h->pat.kind == 0) {   // K2: constraint rows wait for a neighbouring variable (see symbolic.hpp)
        delayed.assign(h->pat.N, 0);
        for (int32_t i = h->pat.n_tot; i < h->pat.N; ++i) delayed[i] = 1;
    }
    MADIPM_TRY(h, analyze(h->pat.N, h->pat.kkt_colptr.data(), h->pat.kkt_rowval.data(), o, h->sym,
                          delayed.empty() ? nullptr : delayed.data()));
    h->analyzed = true;
    return MADIPM_OK;
}

int32_t madipm_get_info(const madipm_handle* h, madipm_info* out) {
    if (!h || !out) return MADIPM_BAD_ARGUMENT;
    std::memset(out, 0, sizeof(*out));
    const KktPattern& p = h->pat;
    out->kkt_kind = h->cfg.kkt_kind == MADIPM_KKT_K2_SCALED ? MADIPM_KKT_K2_SCALED : p.kind;
    out->nvar = p.nx;
    out->n_slack = p.n_slack;
    out->n_tot = p.n_tot;
    out->ncon = p.m;
    out->nlb = h->nlb;
    out->nub = h->nub;
    out->batch = h->cfg.batch;
    out->kkt_order = p.N;
    out->nnzh = p.nnzh;
    out->nnzj = p.nnzj;
    out->kkt_nnz = p.kkt_colptr.empty() ? 0 : p.kkt_colptr.back();
    out->aug_coo_len = static_cast<int64_t>(p.aug_csc_map.size());
    out->n_products = static_cast<int64_t>(p.prod_a.size());
    out->kkt_vec_len = static_cast<int64_t>(p.n_tot) + p.m + h->nlb + h->nub;
    if (h->analyzed) {
        const Symbolic& s = h->sym;
        out->nnz_L = s.nnz_L;
        out->nnz_stored = s.nnz_stored;
        out->factor_flops = s.flops;
        out->nsuper = s.nsuper;
        out->max_height = s.max_h;
        out->max_width = s.max_w;
        out->n_warp_tasks = s.n_warp_tasks;
        out->n_block_tasks = s.n_block_tasks;
        out->n_big = s.n_big;
        out->n_levels = s.n_levels;
    }
    return MADIPM_OK;
}


int32_t madipm_get_symbolic(const madipm_handle* h, int32_t which, int64_t* out, int64_t cap, int64_t* len) {
    if (!h) return MADIPM_BAD_ARGUMENT;
    const KktPattern& p = h->pat;
    const Symbolic& s = h->sym;
    switch (which) {
        case MADIPM_SYM_PERM: return copy_out(s.perm, out, cap, len);
        case MADIPM_SYM_ETREE_PARENT: return copy_out(s.parent, out, cap, len);
        case MADIPM_SYM_COLCOUNT: return copy_out(s.colcount, out, cap, len);
        case MADIPM_SYM_SUPER_FIRST: return copy_out(s.sn_first, out, cap, len);
        case MADIPM_SYM_SUPER_PARENT: return copy_out(s.sn_parent, out, cap, len);
        case MADIPM_SYM_SUPER_ROWPTR: return copy_out(s.sn_rowptr, out, cap, len);
        case MADIPM_SYM_SUPER_ROWS: return copy_out(s.sn_rows, out, cap, len);
        case MADIPM_SYM_SUPER_XOFF: return copy_out(s.sn_xoff, out, cap, len);
        case MADIPM_SYM_ASM_MAP: return copy_out(s.asm_map, out, cap, len);
        case MADIPM_SYM_KKT_COLPTR: return copy_out(p.kkt_colptr, out, cap, len);
        case MADIPM_SYM_KKT_ROWVAL: return copy_out(p.kkt_rowval, out, cap, len);
        case MADIPM_SYM_AUG_CSC_MAP: return copy_out(p.aug_csc_map, out, cap, len);
        case MADIPM_SYM_A_ROWPTR: return copy_out(p.A_rowptr, out, cap, len);
        case MADIPM_SYM_A_COLIND: return copy_out(p.A_colind, out, cap, len);
        case MADIPM_SYM_A_CSR_MAP: return copy_out(p.A_csr_order, out, cap, len);
        case MADIPM_SYM_SUPER_LEVEL: return copy_out(s.sn_level, out, cap, len);
        case MADIPM_SYM_PROD_PTR: return copy_out(p.prod_ptr, out, cap, len);
        case MADIPM_SYM_PROD_A: return copy_out(p.prod_a, out, cap, len);
        case MADIPM_SYM_PROD_B: return copy_out(p.prod_b, out, cap, len);
        case MADIPM_SYM_PROD_K: return copy_out(p.prod_k, out, cap, len);
        case MADIPM_SYM_CHILD_PTR: return copy_out(s.child_ptr, out, cap, len);
        case MADIPM_SYM_CHILD_IDX: return copy_out(s.child_idx, out, cap, len);
        case MADIPM_SYM_SUPER_KIND: return copy_out(s.sn_kind, out, cap, len);
        case MADIPM_SYM_PAR_POS: return copy_out(s.par_pos, out, cap, len);
        case MADIPM_SYM_SUPER_COFF: return copy_out(s.sn_coff, out, cap, len);
        case MADIPM_SYM_TASK_SN: return copy_out(s.task_sn, out, cap, len);
        case MADIPM_SYM_PHASE_PTR: return copy_out(s.phase_ptr, out, cap, len);
        case MADIPM_SYM_PHASE_NBLOCK: return copy_out(s.phase_nblock, out, cap, len);
        case MADIPM_SYM_PHASE_CLS: return copy_out(s.phase_cls, out, cap, len);
        default: return MADIPM_BAD_ARGUMENT;
    }
}

int64_t madipm_launch_count(const madipm_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
