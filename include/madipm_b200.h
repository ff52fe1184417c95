/*
 * madipm_b200.h -- C ABI of libmadipm_b200.so, the B200 (sm_100a) backend for the
 * per-iteration KKT hot path of MadIPM / MadQP.jl.
 *
 * This is what a Julia `ccall` (or any FFI) binds.  Plain pointers and sizes only.
 * Index arrays that describe the problem are 1-based exactly as the Julia caller
 * holds them (Int32 COO triplets, /root/reference/src/KKT/normalkkt.jl:51-77;
 * Vector{Int} index sets, /root/reference/src/structure.jl:95-102); everything the
 * library hands back through the symbolic getters is 0-based.
 *
 * Ownership: the caller owns every host array; the library copies during the call
 * and never keeps a host pointer.  All device memory belongs to the handle.
 * Threading: one handle = one CUDA stream = one caller thread; not re-entrant.
 * Calls are stream-ordered; only calls that return scalars synchronise.
 * Errors: no exceptions cross the boundary; every function returns a status code.
 *
 * There is no CPU fallback: numeric entry points return MADIPM_CUDA_ERROR when no
 * sm_100-class device is present.  The symbolic entry points (set_pattern_*,
 * analyze, the getters) are host code and work without a GPU.
 */
#ifndef MADIPM_B200_H
#define MADIPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct madipm_handle madipm_handle;

/* ---- status codes (SURVEY.md 8b "Errors") ------------------------------------------ */
enum {
    MADIPM_OK = 0,
    MADIPM_FACTOR_BREAKDOWN = 1, /* zero / non-finite pivot: feeds is_factorized(), /root/reference/src/utils.jl:54-62 */
    MADIPM_NONFINITE = 2,        /* NaN in a solve residual: SolveException, /root/reference/src/linear_solver.jl:41-43 */
    MADIPM_CUDA_ERROR = 3,       /* LinearSolverException on the Julia side, /root/reference/src/solver.jl:381-383 */
    MADIPM_BAD_ARGUMENT = 4,
    MADIPM_INTERNAL = 5
};

/* ---- configuration ------------------------------------------------------------------- */
enum { MADIPM_KKT_K2 = 0, MADIPM_KKT_NORMAL = 1, MADIPM_KKT_K2_SCALED = 2 };
enum { MADIPM_ORDER_AMD = 0, MADIPM_ORDER_NATURAL = 1, MADIPM_ORDER_ND = 2 /* nested dissection, AMD leaves (default) */ };

typedef struct {
    int32_t kkt_kind;       /* MADIPM_KKT_K2: MadNLP.SparseKKTSystem (default, /root/reference/src/utils.jl:108);
                               MADIPM_KKT_NORMAL: MadIPM.NormalKKTSystem (/root/reference/src/KKT/normalkkt.jl);
                               MADIPM_KKT_K2_SCALED: MadNLP.ScaledSparseKKTSystem ("K2.5": the K2 system scaled
                               symmetrically by sqrt of the bound slacks; /root/reference/src/kernels.jl:149-165,
                               /root/reference/scripts/cuda_wrapper.jl:90-116, /root/reference/test/runtests.jl:102-115).
                               Same pattern entry point, symbolic analysis and factor kernels as MADIPM_KKT_K2 */
    int32_t ordering;
    int32_t max_width;      /* supernode width cap */
    int32_t relax_small;    /* relaxed amalgamation, see csrc/symbolic.hpp */
    double relax_zeros[3];
    int64_t block_panel;    /* largest panel (doubles) one thread block factorises in shared memory; larger
                               supernodes take the multi-block (DMMA) path.  See csrc/symbolic.hpp */
    int32_t device;         /* CUDA device ordinal */
    int32_t batch;          /* instances sharing one pattern (config C5); 1 for a single solve */
    int32_t nd_leaf;        /* nested dissection leaf size, 0 = automatic */
    int32_t reserved;
} madipm_config;

void madipm_default_config(madipm_config* cfg);

/* ---- lifecycle ------------------------------------------------------------------------- */
int32_t madipm_create(const madipm_config* cfg, madipm_handle** out);
int32_t madipm_destroy(madipm_handle* h);
const char* madipm_last_error(const madipm_handle* h);
const char* madipm_version(void);

/* ---- symbolic phase, once per pattern (replaces create_kkt_system + the linear-solver
 *      constructor, /root/reference/src/KKT/normalkkt.jl:29-126) ------- */
int32_t madipm_set_pattern_k2(madipm_handle* h, int32_t nvar, int32_t ncon,
                              int64_t nnzh, const int32_t* hess_I, const int32_t* hess_J,
                              int64_t nnzj, const int32_t* jac_I, const int32_t* jac_J,
                              int32_t n_slack, const int64_t* ind_ineq);
int32_t madipm_set_pattern_normal(madipm_handle* h, int32_t nvar, int32_t ncon,
                                  int64_t nnzj, const int32_t* jac_I, const int32_t* jac_J,
                                  int32_t n_slack, const int64_t* ind_ineq);
/* ind_lb / ind_ub index into [x ; s] (/root/reference/src/structure.jl:144-151) */
int32_t madipm_set_bounds_index(madipm_handle* h, int32_t nlb, const int64_t* ind_lb,
                                int32_t nub, const int64_t* ind_ub);
int32_t madipm_analyze(madipm_handle* h);

typedef struct {
    int32_t kkt_kind, nvar, n_slack, n_tot, ncon, nlb, nub, batch;
    int32_t kkt_order;       /* N: order of the matrix that is factorised */
    int64_t nnzh, nnzj;
    int64_t kkt_nnz;         /* nnz of the lower-triangular CSC handed to the factorisation */
    int64_t aug_coo_len;     /* K2: length of the augmented COO value array */
    int64_t nnz_L;           /* exact nnz(L) */
    int64_t nnz_stored;      /* doubles of supernodal panel storage */
    double factor_flops;     /* sum_j colcount_j^2 */
    int32_t nsuper, max_height, max_width;
    int32_t n_warp_tasks;    /* supernodes handled by one warp */
    int32_t n_levels;        /* heights of the supernodal elimination tree = phases of the factor / solve kernels */
    int64_t n_products;      /* normal equations: length of the product lists */
    int64_t kkt_vec_len;     /* n_tot + ncon + nlb + nub: length of an UnreducedKKTVector */
    int32_t n_block_tasks;   /* supernodes handled by one thread block */
    int32_t n_big;           /* supernodes on the multi-block path */
} madipm_info;
int32_t madipm_get_info(const madipm_handle* h, madipm_info* out);

/* integer results of the symbolic phase, copied to `out` (capacity `cap` entries);
 * *len receives the length.  All 0-based. */
enum {
    MADIPM_SYM_PERM = 0,          /* perm[new] = old, length N */
    MADIPM_SYM_ETREE_PARENT = 1,  /* parent in the permuted numbering, -1 root */
    MADIPM_SYM_COLCOUNT = 2,
    MADIPM_SYM_SUPER_FIRST = 3,   /* nsuper + 1 */
    MADIPM_SYM_SUPER_PARENT = 4,
    MADIPM_SYM_SUPER_ROWPTR = 5,  /* nsuper + 1 */
    MADIPM_SYM_SUPER_ROWS = 6,
    MADIPM_SYM_SUPER_XOFF = 7,    /* nsuper + 1 */
    MADIPM_SYM_ASM_MAP = 8,       /* kkt_nnz: CSC nz -> panel slot */
    MADIPM_SYM_KKT_COLPTR = 9,    /* N + 1 */
    MADIPM_SYM_KKT_ROWVAL = 10,   /* kkt_nnz */
    MADIPM_SYM_AUG_CSC_MAP = 11,  /* K2: aug COO entry -> CSC nz */
    MADIPM_SYM_A_ROWPTR = 12,     /* CSR of A incl. slack columns */
    MADIPM_SYM_A_COLIND = 13,
    MADIPM_SYM_A_CSR_MAP = 14,    /* CSR slot -> COO index (0-based A_csr_map, normalkkt.jl:83-88) */
    MADIPM_SYM_SUPER_LEVEL = 15,  /* height of the supernode in the supernodal elimination tree (0 = leaf) */
    MADIPM_SYM_PROD_PTR = 16,     /* normal equations product lists */
    MADIPM_SYM_PROD_A = 17,
    MADIPM_SYM_PROD_B = 18,
    MADIPM_SYM_PROD_K = 19,
    MADIPM_SYM_CHILD_PTR = 20,    /* nsuper + 1: children lists of the supernodal elimination tree */
    MADIPM_SYM_CHILD_IDX = 21,
    MADIPM_SYM_SUPER_KIND = 22,   /* 0 small (parent pulls its contribution block), 1 small under a big parent, 2 big */
    MADIPM_SYM_PAR_POS = 23,      /* like SUPER_ROWS: position of a below-row in the parent's row list */
    MADIPM_SYM_SUPER_COFF = 24,   /* nsuper + 1: offsets of the packed contribution blocks */
    MADIPM_SYM_TASK_SN = 25,      /* small supernodes in schedule order */
    MADIPM_SYM_PHASE_PTR = 26,    /* n_levels + 1 into TASK_SN */
    MADIPM_SYM_PHASE_NBLOCK = 27, /* n_levels: tasks of a phase that are not warp tasks */
    MADIPM_SYM_PHASE_CLS = 28     /* n_levels x 5: first task of class 256 / 128 / 64 / 32 (threads per task), end of the phase */
};
int32_t madipm_get_symbolic(const madipm_handle* h, int32_t which, int64_t* out, int64_t cap, int64_t* len);

/* ---- values, once per solve (eval_jac_wrapper! / eval_lag_hess_wrapper! +
 *      compress_jacobian! / compress_hessian!, /root/reference/src/solver.jl:166-170,
 *      /root/reference/src/KKT/normalkkt.jl:149-160).  Arrays are batch-major. ----------- */
int32_t madipm_set_values(madipm_handle* h, const double* hess_V, const double* jac_V);

/* One call for everything solve! does with the problem data before the first iteration
 * (/root/reference/src/solver.jl:131-170): MadNLP.initialize!(cb, x, xl, xu, y, rhs, ind_ineq; tol, bound_push,
 * bound_fac) -- bounds relaxed by tol, x0 pushed into their interior, slacks started at the pushed constraint values --,
 * MadNLP.set_scaling! (con_scale = min(1, max_gradient / row max of |J|), obj_scale = min(1, max_gradient / max |grad f(x0)|)),
 * and the evaluation + compression of the scaled Jacobian / Hessian values (what madipm_set_values does).
 * Inputs are HOST arrays of the unscaled model, batch-major: x0, lvar, uvar, c [nvar]; y0, lcon, ucon [ncon];
 * hess_V [nnzh], jac_V [nnzj].  All of it runs on the device.  Outputs (host, one per instance): obj_scale and
 * norm_b = ||rhs||_inf (solver.jl:173); con_scale stays on the device as MADIPM_VEC_CON_SCALE. */
typedef struct {
    double tol;            /* bound_relax_factor */
    double bound_push, bound_fac;
    double max_gradient;   /* 100 in the reference (solver.jl:148-159) */
    int32_t scaling;       /* 0: con_scale = obj_scale = 1 */
    int32_t reserved;
} madipm_load_options;
int32_t madipm_load_problem(madipm_handle* h, const double* x0, const double* lvar, const double* uvar, const double* y0,
                            const double* lcon, const double* ucon, const double* c, const double* hess_V, const double* jac_V,
                            const madipm_load_options* opt, double* obj_scale, double* norm_b);

/* ---- device vectors owned by the handle (MPCSolver fields, /root/reference/src/structure.jl:1-75,
 *      and the KKT diagonals, /root/reference/src/KKT/normalkkt.jl:10-18) ------------------ */
enum {
    MADIPM_VEC_X = 0, MADIPM_VEC_XL, MADIPM_VEC_XU, MADIPM_VEC_ZL, MADIPM_VEC_ZU,   /* n_tot */
    MADIPM_VEC_F, MADIPM_VEC_JACL, MADIPM_VEC_COST,                                 /* n_tot */
    MADIPM_VEC_Y, MADIPM_VEC_C, MADIPM_VEC_RHS,                                     /* ncon */
    MADIPM_VEC_D, MADIPM_VEC_P, MADIPM_VEC_W1, MADIPM_VEC_W2,                       /* kkt_vec_len */
    MADIPM_VEC_CORR_LB, MADIPM_VEC_CORR_UB,                                         /* nlb, nub */
    MADIPM_VEC_REG, MADIPM_VEC_PR_DIAG,                                             /* n_tot */
    MADIPM_VEC_DU_DIAG,                                                             /* ncon */
    MADIPM_VEC_L_DIAG, MADIPM_VEC_L_LOWER,                                          /* nlb */
    MADIPM_VEC_U_DIAG, MADIPM_VEC_U_LOWER,                                          /* nub */
    MADIPM_VEC_KKT_NZVAL,                                                           /* kkt_nnz (aug_com.nzval) */
    MADIPM_VEC_CON_SCALE,                                                           /* ncon (set_scaling!, see madipm_load_problem) */
    MADIPM_VEC_SCALING,                                                             /* n_tot (K2.5 kkt.scaling_factor; ones otherwise) */
    MADIPM_VEC_COUNT
};
int32_t madipm_vec_len(const madipm_handle* h, int32_t vec, int64_t* len);     /* per instance */
int32_t madipm_vec_set(madipm_handle* h, int32_t vec, const double* host);     /* batch * len doubles */
int32_t madipm_vec_get(madipm_handle* h, int32_t vec, double* host);
int32_t madipm_vec_fill(madipm_handle* h, int32_t vec, double value);
int32_t madipm_vec_device_ptr(madipm_handle* h, int32_t vec, void** dev);      /* for zero-copy interop */
int32_t madipm_synchronize(madipm_handle* h);
void* madipm_stream(madipm_handle* h);                                         /* cudaStream_t */

/* ---- AbstractKKTSystem / linear-solver operations --------------------------------------- */
/* MadNLP.initialize!(kkt), /root/reference/src/KKT/normalkkt.jl:136-147 */
int32_t madipm_kkt_initialize(madipm_handle* h);
/* K1 set_aug_diagonal_reg!, /root/reference/src/kernels.jl:128-146 (one fused kernel).
 * MADIPM_KKT_K2_SCALED: the ScaledSparseKKTSystem method, /root/reference/src/kernels.jl:149-165 (l_diag = x - xl,
 * u_diag = xu - x, both positive) fused with MadNLP._set_aug_diagonal! (scaling_factor, pr_diag) */
int32_t madipm_set_aug_diagonal_reg(madipm_handle* h, double del_w, double del_c);
/* solver.jl:16-18: reg = pr_diag = del_w, du_diag = del_c */
int32_t madipm_set_initial_regularization(madipm_handle* h, double del_w, double del_c);
/* K2a / K2b build_kkt!: K2 transfer! (/root/reference/scripts/cuda_wrapper.jl:17-28) or
 * normal-equations assembly (/root/reference/src/utils.jl:266-298) into aug_com.nzval.
 * MADIPM_KKT_K2_SCALED: _build_scale_augmented_system_coo! (/root/reference/scripts/cuda_wrapper.jl:90-116)
 * + transfer!, one pass over the nonzeros of aug_com */
int32_t madipm_build_kkt(madipm_handle* h);
/* factorize!: supernodal LDL' (K2) / Cholesky-as-LDL' (normal).  *n_failed = instances whose
 * factorisation broke down; status MADIPM_FACTOR_BREAKDOWN when > 0. Synchronises. */
int32_t madipm_factorize(madipm_handle* h, int32_t* n_failed);
/* (n_failed == NULL: the call does not synchronise; the breakdown flags are collected by the next call that
 *  synchronises anyway -- any call returning scalars, madipm_is_factorized, madipm_synchronize -- and the status is
 *  then read with madipm_is_factorized.  The retry of linear_solver.jl:8-16 still sees every failure, one call later.) */
/* is_factorized per instance (0/1), /root/reference/src/utils.jl:54-62 */
int32_t madipm_is_factorized(madipm_handle* h, int32_t* flags);
/* The LDL' does not pivot (neither does the reference's default CPU solver, LDLFactorizations).  A pivot of roundoff size --
 * linearly dependent constraint rows with delta_c = 0 -- is not exactly zero, passes the zero / non-finite test and yields a
 * garbage factor.  Two more breakdown tests catch it: |pivot| <= tiny (default 1e-13 in the scaled problem's units: the
 * regularised pivots of an interior-point KKT matrix are 1e-8 and up, cancellation leaves 1e-16 times the operands), and
 * element growth, a factor entry |L_ik| above `growth` (default 1e13: regularisation 1e-8 against entries of order one
 * explains 1e8).  0 switches a test off; (0, 0) is the reference's exact semantics and the default for the normal
 * equations, whose pivots and factor entries legitimately span many orders of magnitude (Sigma does). */
int32_t madipm_set_pivot_checks(madipm_handle* h, double tiny, double growth);
/* Scheduling of the small-front tree kernels (the numeric factorisation K3 and the triangular sweeps K4, no reference
 * counterpart: LDLFactorizations is sequential).  0: a grid-wide barrier between two heights of the supernodal elimination
 * tree; 1: dataflow -- every task waits for its own children (parent in the backward sweep) through per-task counters in
 * global memory, no barrier between heights.  Same tasks, same arithmetic, bit-identical results.  Only for workloads whose
 * fronts all fit one block (no multi-block fronts); *in_effect (may be NULL) says whether the launches will use it.
 * The triangular sweeps keep the barriers where every height of the launch has at least two tasks per block (large
 * batches: throughput-bound, the per-task signalling would cost more than the barriers).
 * Default: on; environment variable MADIPM_TREE_DATAFLOW=0 switches it off. */
int32_t madipm_set_tree_dataflow(madipm_handle* h, int32_t on, int32_t* in_effect);
/* inertia of the last factorisation of instance `inst`: signs of D */
int32_t madipm_inertia(madipm_handle* h, int32_t inst, int64_t* n_pos, int64_t* n_zero, int64_t* n_neg);
/* solve!(linear_solver, rhs): in-place solve with the current factors on a device vector of
 * length N per instance (pointer obtained from madipm_vec_device_ptr + offset) */
int32_t madipm_linear_solve(madipm_handle* h, double* dev_rhs, int64_t stride);
/* K4 / K4' MadNLP.solve!(kkt, w) on a KKT vector (reduce_rhs!, solve, finish_aug_solve!),
 * /root/reference/src/KKT/normalkkt.jl:182-205 */
int32_t madipm_kkt_solve(madipm_handle* h, int32_t vec_w);
/* K5 mul!(w, kkt, v, alpha, beta), /root/reference/src/KKT/normalkkt.jl:207-219 */
int32_t madipm_kkt_mul(madipm_handle* h, int32_t vec_w, int32_t vec_v, double alpha, double beta);
/* K6 solve_system!(d, solver, p), /root/reference/src/linear_solver.jl:19-45.
 * norms[3 * inst + {0,1,2}] = ||p - K d||inf, ||p||inf, ||d||inf.  Synchronises.
 * Returns MADIPM_NONFINITE if any residual ratio is NaN. */
int32_t madipm_solve_system(madipm_handle* h, int32_t vec_d, int32_t vec_p, double* norms);
/* K7 jtprod!: out(n_tot) = A' * in(ncon); and the model products of K8 */
enum { MADIPM_OP_A = 0, MADIPM_OP_AT = 1, MADIPM_OP_Q = 2 };
int32_t madipm_spmv(madipm_handle* h, int32_t op, int32_t vec_out, int32_t vec_in);
/* same on raw device pointers (per-instance strides in doubles) */
int32_t madipm_spmv_ptr(madipm_handle* h, int32_t op, double* dev_out, int64_t stride_out,
                        const double* dev_in, int64_t stride_in);

/* ---- fused phases of mpc!.  Scalars come back per instance. ------------ */
/* K8: f = Qx + cost, c = A x - s - rhs, obj = c0 + cost'x + x'Qx/2  (solver.jl:338-340).
 * obj[inst]; does not synchronise unless obj != NULL. */
int32_t madipm_eval_model(madipm_handle* h, double c0, double* obj);
/* K7 + K15: jacl = A'y, then out[4*inst + ..] = { ||c||inf, ||f - zl + zu + jacl||inf,
 * max complementarity (mu = 0), ||dx||inf }   (solver.jl:259-274) */
int32_t madipm_phase_residuals(madipm_handle* h, double* out);
/* K9 set_initial_primal_rhs! / set_initial_dual_rhs! (kernels.jl:1-19) */
int32_t madipm_set_initial_primal_rhs(madipm_handle* h);
int32_t madipm_set_initial_dual_rhs(madipm_handle* h);
/* K9 set_predictive_rhs! (kernels.jl:21-41) */
int32_t madipm_set_predictive_rhs(madipm_handle* h);
/* K9 set_correction_rhs! (kernels.jl:43-61); mu per instance */
int32_t madipm_set_correction_rhs(madipm_handle* h, const double* mu);
/* K12 get_fraction_to_boundary_step (kernels.jl:242-305): out[8*inst + ..] =
 * { alpha_xl, alpha_xu, alpha_zl, alpha_zu, idx_xl, idx_xu, idx_zl, idx_zu } (indices 1-based, 0 = none) */
int32_t madipm_ratio_test(madipm_handle* h, const double* tau, double* out);
/* K11 get_complementarity_measure / get_affine_complementarity_measure (kernels.jl:171-224);
 * alpha_p / alpha_d per instance (NULL = 0: current complementarity) */
int32_t madipm_complementarity(madipm_handle* h, const double* alpha_p, const double* alpha_d, double* out);
/* K12 + K11 + K10 fused, the predictor bookkeeping of mpc! (solver.jl:295-297):
 * ratio test with tau = 1, affine complementarity at those steps, corrections, current mu.
 * out[4*inst + ..] = { alpha_aff_p, alpha_aff_d, mu_affine, mu_curr } */
int32_t madipm_phase_affine(madipm_handle* h, double* out);
/* K10 get_correction! (kernels.jl:63-75) and Gondzio's set_extra_correction! (kernels.jl:78-126) */
int32_t madipm_get_correction(madipm_handle* h);
int32_t madipm_set_extra_correction(madipm_handle* h, const double* alpha_p, const double* alpha_d,
                                    double beta_min, double beta_max, const double* mu);
/* K14 iterate update + adjust_boundary! (solver.jl:332-335, 342); `active` (NULL = all) masks
 * finished instances of a batch */
int32_t madipm_update_iterate(madipm_handle* h, const double* alpha_p, const double* alpha_d,
                              const double* mu, const int32_t* active);
/* MehrotraAdaptiveStep gathers (kernels.jl:351-368) without scalar device reads:
 * out[inst*2 + ..] = { alpha_p, alpha_d } for gamma_f */
int32_t madipm_mehrotra_step(madipm_handle* h, double gamma_f, double* out);
/* starting point, solver.jl:35-118 (steps 3 onwards, after the two least-squares solves) */
int32_t madipm_init_multipliers_and_shift(madipm_handle* h, double bound_fac);
/* vector helpers the control loop uses directly */
int32_t madipm_vec_copy(madipm_handle* h, int32_t dst, int32_t src);
/* dst[0:len) += alpha * src[off_src : off_src+len) at offset off_dst (axpy on sub-ranges) */
int32_t madipm_vec_axpy(madipm_handle* h, int32_t dst, int64_t off_dst, int32_t src, int64_t off_src,
                        int64_t len, double alpha);
int32_t madipm_vec_norm_inf(madipm_handle* h, int32_t vec, double* out);

/* ---- sub-range and index-view operations --------------------------------------------------------------------
 * What the Julia vector type of julia/MadIPMB200.jl (tag + handle + range) maps its methods to, so that the lines of
 * /root/reference/src/solver.jl that work on solver vectors directly (axpy! at :28 and :332-333, `.=` at :16-18 and :33,
 * the `.+=` on the zl_r / zu_r index views at :334-335, copyto! of linear_solver.jl:25-29 and solver.jl:236-242,
 * norm(., Inf) at solver.jl:174-175 and structure.jl:191, adjust_boundary! at :342) run unchanged.  Ranges are 0-based
 * [off, off + len) inside one instance and apply to every instance of a batch. */
int32_t madipm_vec_get_range(madipm_handle* h, int32_t vec, int64_t off, int64_t len, double* host);       /* host: [batch][len] */
int32_t madipm_vec_set_range(madipm_handle* h, int32_t vec, int64_t off, int64_t len, const double* host);
int32_t madipm_vec_fill_range(madipm_handle* h, int32_t vec, int64_t off, int64_t len, double value);
int32_t madipm_vec_copy_range(madipm_handle* h, int32_t dst, int64_t off_dst, int32_t src, int64_t off_src, int64_t len);
int32_t madipm_vec_norm_inf_range(madipm_handle* h, int32_t vec, int64_t off, int64_t len, double* out);   /* out: [batch] */
/* dst[ind[k]] += alpha * src[off_src + k], k < nlb (which = 0: ind = ind_lb) or k < nub (which = 1: ind = ind_ub);
 * dst is one of the n_tot-long vectors.  solver.jl:334-335: zl_r .+= alpha_d .* dual_lb(d) */
int32_t madipm_vec_axpy_at(madipm_handle* h, int32_t dst, int32_t which, int32_t src, int64_t off_src, double alpha);
/* MadNLP.adjust_boundary!(x_lr, xl_r, x_ur, xu_r, mu), solver.jl:342 ([RECALLED], same arithmetic as inside
 * madipm_update_iterate); mu: one value per instance */
int32_t madipm_adjust_boundary(madipm_handle* h, const double* mu);

/* ---- device-resident inputs ------------------------------------------------------------------
 * A solve overwrites x, y and (through adjust_boundary!) xl, xu.  snapshot keeps device copies of the
 * four input vectors as uploaded; restore puts them back without host traffic, so a solve can be
 * repeated with its inputs already resident in HBM (bench.py's device-resident figure). */
int32_t madipm_snapshot_inputs(madipm_handle* h);
int32_t madipm_restore_inputs(madipm_handle* h);

/* ---- page-locked host memory -------------------------------------------------------------------------------
 * Host arrays stay the caller's (header, "Ownership"), and any host memory works.  Copies from / to PAGE-LOCKED memory are
 * the fast and truly asynchronous ones (a copy from pageable memory synchronises the stream and goes through a staging
 * buffer at a fraction of the PCIe rate), so a caller that re-sends inputs every solve keeps them in memory from this
 * allocator (the Julia side: the buffers behind the model's x0, lvar, ..., the Jacobian / Hessian values).  Needs no handle. */
int32_t madipm_pinned_alloc(int64_t bytes, void** ptr);
int32_t madipm_pinned_free(void* ptr);

/* ---- measurement hooks -------------------------------------------------------------------- */
/* number of kernel launches issued by this handle since creation */
int64_t madipm_launch_count(const madipm_handle* h);
/* accumulated device milliseconds per kernel class since the last reset (CUDA events on the
 * handle's stream; enabled by madipm_profile(h, 1)) */
enum { MADIPM_T_ASSEMBLY = 0, MADIPM_T_FACTOR, MADIPM_T_SOLVE, MADIPM_T_SPMV, MADIPM_T_VECTOR,
       MADIPM_T_FACTOR_TREE /* the tree launches of the first segment alone */,
       MADIPM_T_FACTOR_TOP /* ... of which the launch for the heights near the root (k_front / k_tree_top) */, MADIPM_T_COUNT };
/* enable = 1: every class, with a stream synchronisation per call (diagnostic runs only);
 * enable = 2: only the factorisation (classes FACTOR and FACTOR_TREE), read at the synchronisation
 *             madipm_factorize performs anyway -- cheap enough to leave on inside a timed region. */
int32_t madipm_profile(madipm_handle* h, int32_t enable);
int32_t madipm_profile_read(madipm_handle* h, double* ms, int64_t* calls, int32_t reset);

/* ---- hang guards (the reference has none: it never spins on the device) --------------------------------------------
 * The persistent tree kernels synchronise the grid with a counter in global memory.  That spin is bounded
 * (MADIPM_BARRIER_TIMEOUT_MS, default 2000): a block that runs out of patience raises a flag that makes the whole grid
 * leave, and the next synchronising entry point returns MADIPM_INTERNAL with the barrier's state in madipm_last_error.
 * Every host-side stream synchronisation is bounded as well (MADIPM_SYNC_TIMEOUT_S, default 120; 0 = unbounded); on
 * expiry the call returns MADIPM_INTERNAL and the handle is wedged (every later numeric call fails at once).
 * madipm_debug_barrier_state: copy of the device-written record (host-visible memory: readable while a kernel is stuck);
 * out8 = {aborted, kernel, op, height, block, grid, seen, target}.
 * madipm_debug_inject_barrier_fault: test hook -- the next tree launch starts with a barrier counter that can never
 * reach its target, so the guards above can be exercised on purpose (tests/test_gpu_guards.py). */
int32_t madipm_debug_barrier_state(madipm_handle* h, int64_t* out8);
int32_t madipm_debug_inject_barrier_fault(madipm_handle* h);
/* One line of text: the barrier record, the kernel launched last, the launch count and the last height a tree kernel
 * started.  Makes no CUDA call and takes no lock, so a watchdog thread may call it while the owning thread is blocked. */
const char* madipm_debug_describe(madipm_handle* h);

/* ---- fused iteration phases: one host synchronisation each ------------------------------------------ */
/* The reference synchronises once per reduction (~22 per iteration); the separate entry points above once per returned
 * group of scalars (7 per iteration); these three bring an iteration of mpc! (solver.jl:254-345) to 3. */
/* solver.jl:259-274 + the objective left on the device by the last madipm_eval_model(h, c0, NULL) (solver.jl:338) */
int32_t madipm_phase_residuals_obj(madipm_handle* h, double c0, double* out4, double* obj);
/* affine_direction!, solver.jl:294-297: set_predictive_rhs! + solve_system!(d, solver, p) + ratio tests at tau = 1 +
 * affine complementarity + get_correction!.  norms: 3 per instance as madipm_solve_system, out: 4 per instance as
 * madipm_phase_affine.  MADIPM_NONFINITE as madipm_solve_system. */
int32_t madipm_phase_predictor(madipm_handle* h, double* norms, double* out4);
/* The same preceded by the first try of factorize_regularized_system! (linear_solver.jl:6-17: set_aug_diagonal_reg! +
 * build_kkt! + factorize!), one synchronisation for all of it.  When every supernode is factorised by the tree kernels the
 * predictor's forward sweep is done INSIDE the factor launch (its right-hand side depends on the iterate only,
 * kernels.jl:21-41; each supernode's forward step runs while its freshly factorised panel is still in shared memory): an
 * iteration then costs three triangular sweeps instead of four.  Same arithmetic as the separate calls, identical results.
 * Breakdown flags as after madipm_factorize(h, NULL) (madipm_is_factorized); on a breakdown the caller re-enters the
 * reference's retry loop with the separate entry points. */
int32_t madipm_phase_factor_predictor(madipm_handle* h, double del_w, double del_c, double* norms, double* out4);
/* mpc_direction! without Gondzio corrections, solver.jl:307-311, + the ratio tests of update_step! (kernels.jl:290-305):
 * set_correction_rhs!(mu) + solve_system! + ratio tests at tau.  ratio: 8 per instance as madipm_ratio_test. */
int32_t madipm_phase_corrector(madipm_handle* h, const double* mu, const double* tau, double* norms, double* ratio8);

/* One whole iteration of mpc! (solver.jl:288-342) with ONE synchronisation, for max_ncorr = 0 and the Conservative / Adaptive
 * step rules: madipm_phase_factor_predictor, update_barrier! (kernels.jl:226-236), madipm_phase_corrector, update_step!,
 * madipm_update_iterate, madipm_eval_model and the residuals of the NEW iterate (what the next pass of the loop starts
 * with).  The scalars the host computes between the phases (mu_affine, mu_curr, sigma, mu, tau, alpha_p, alpha_d) are
 * computed on the device in the same floating-point operations.  The iterate is updated only if, for every active
 * instance, the factorisation did not break down and both residual ratios are numbers; otherwise *ok = 0, nothing has
 * moved, and the caller redoes the iteration through the separate phases (the reference's retry, linear_solver.jl:8-16).
 * active: one flag per instance or NULL (all).  out: 28 doubles per instance:
 *   [0:3) predictor norms (as madipm_solve_system), [3:7) alpha_aff_p, alpha_aff_d, mu_affine, mu_curr, [7] mu, [8] tau,
 *   [9:12) corrector norms, [12:20) ratio tests at tau (as madipm_ratio_test), [20:22) alpha_p, alpha_d,
 *   [22:26) residuals of the new iterate (as madipm_phase_residuals), [26] its objective c0 + c'x + x'Qx/2, [27] unused. */
typedef struct {
    double del_w, del_c;      /* regularisation of this iteration */
    double mu_min;            /* opt.mu_min */
    double tau;               /* ConservativeStep: tau;  AdaptiveStep: tau_min */
    double c0;                /* constant term of the objective (scaled) */
    double res_tol;           /* > 0: a predictor residual ratio above it is a no-go as well (a garbage factor that the pivot tests
                                 cannot see: exactly dependent rows); 0: only NaN is, as in linear_solver.jl:41-43 */
    int32_t adaptive_step;    /* 1: tau = max(1 - mu, tau_min) (kernels.jl:315-323), 0: tau fixed (kernels.jl:307-313) */
    int32_t has_inequalities; /* SURVEY 8a note B.1: any finite bound */
} madipm_iter_options;
int32_t madipm_phase_iteration(madipm_handle* h, const madipm_iter_options* opt, const int32_t* active, double* out28, int32_t* ok);

/* ---- preprocessing --------------------------------------------------------------- */
/* scale_qp, /root/reference/scripts/common.jl:38-100: Ruiz equilibration of the constraint matrix in the infinity
 * norm, As = A ./ (Dr_i * Dc_j) with row and column norms close to 1, and the matching rescaling of the model
 * (H ./ (Dc_i*Dc_j), c ./ Dc, lvar/uvar/x0 .* Dc, lcon/ucon/y0 ./ Dr).  The reference takes (Dr, Dc) from
 * HSL.mc77(A, 0) (proprietary, un-vendored); this restates the published simultaneous row/column iteration (Ruiz,
 * RAL-TR-2001-034).  Needs no handle.  1-based COO triplets; every value array is rescaled IN PLACE on the host side
 * (the work is done on the device: upload, max_iter sweeps, rescale, download); any of c, lvar, uvar, lcon, ucon, x0,
 * y0, Dr, Dc may be NULL.  `tol <= 0`: all max_iter sweeps; otherwise stop once max |1 - norm| <= tol.
 * *final_err = that deviation as seen by the last sweep.  opt->launches returns the number of kernels launched. */
typedef struct {
    int32_t device;
    int32_t max_iter;   /* default 10 */
    double tol;         /* default 0 */
    int64_t launches;   /* out */
} madipm_scale_options;
void madipm_default_scale_options(madipm_scale_options* opt);
int32_t madipm_scale_qp(madipm_scale_options* opt, int32_t nvar, int32_t ncon,
                        int64_t nnzj, const int32_t* jac_I, const int32_t* jac_J, double* jac_V,
                        int64_t nnzh, const int32_t* hess_I, const int32_t* hess_J, double* hess_V,
                        double* c, double* lvar, double* uvar, double* lcon, double* ucon, double* x0, double* y0,
                        double* Dr, double* Dc, int32_t* sweeps_done, double* final_err);
const char* madipm_scale_last_error(void);   /* of the calling thread's last madipm_scale_qp / madipm_standard_form_qp */

/* standard_form_qp, /root/reference/scripts/common.jl:156-288: slack variables s = A x for the inequality rows
 * (lcon < ucon, common.jl:164-168; appended columns nvar+1 .. nvar+ns, entries -1, common.jl:214-219) and every RANGE
 * bound of x or s (-Inf < l < u < Inf, common.jl:180-183, 192-195) moved into an equality row  x_i + w_k = u_i  with a
 * new nonnegative variable w_k (rows ncon+1 .. ncon+nw, columns nvar+ns+1 .., common.jl:221-230, 249-259).  Variables
 * with only an upper bound keep it, fixed variables stay fixed (common.jl:257-259), the Hessian is unchanged (only its
 * dimension grows, common.jl:209).  Needs no handle; the work is done on the device (classification of the n + m
 * bounds, one prefix sum for the ranks k, ranked writes), host buffers in and out.  1-based COO triplets.
 * Outputs, caller-allocated for the worst case, of which the first ... entries are written:
 *   jac_*_out  [nnzj + ncon + 2 (nvar + ncon)]   nnzj + ns + 2 nw   ([A | slack | range pairs], the reference's order)
 *   c_out, lvar_out, uvar_out, x0_out [2 (nvar + ncon)]   nvar + ns + nw
 *   lcon_out, ucon_out, y0_out [nvar + 2 ncon]            ncon + nw
 * c, x0, y0 may be NULL (zeros); any output pointer may be NULL (not downloaded).  nvar + ncon < 2^30. */
int32_t madipm_standard_form_qp(int32_t device, int32_t nvar, int32_t ncon, int64_t nnzj, const int32_t* jac_I, const int32_t* jac_J,
                                const double* jac_V, const double* c, const double* lvar, const double* uvar, const double* lcon,
                                const double* ucon, const double* x0, const double* y0, int32_t* ns, int32_t* nw,
                                int32_t* jac_I_out, int32_t* jac_J_out, double* jac_V_out, double* c_out, double* lvar_out,
                                double* uvar_out, double* x0_out, double* lcon_out, double* ucon_out, double* y0_out,
                                int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* MADIPM_B200_H */
