// Host-side symbolic analysis for the MadIPM KKT hot path (done once per
// sparsity pattern, uploaded to the GPU afterwards).  Reference call stack:
// SURVEY.md 3.1 -- /root/reference/src/KKT/normalkkt.jl:29-126 (normal
// equations), MadNLP SparseKKTSystem (K2, un-vendored) and the linear-solver
// constructor at normalkkt.jl:99-101.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace madipm {

// ---- S1: stable counting sort COO -> CSR.  /root/reference/src/utils.jl:148-191
// 0-based in/out.  order[k] = COO index stored at CSR slot k  (== A_csr_map - 1,
// /root/reference/src/KKT/normalkkt.jl:83-88).
void coo_to_csr(int32_t n_rows, int64_t nnz, const int32_t* Ai, const int32_t* Aj,
                std::vector<int64_t>& Bp, std::vector<int32_t>& Bj, std::vector<int64_t>& order);

// ---- S3: pattern of tril(A A').  Same output as /root/reference/src/utils.jl:199-264
// (column i holds rows j >= i sharing a column index with row i, ascending), but
// by a column-list algorithm instead of the reference's O(m * nnz) scan.
void build_normal_pattern(int32_t n_rows, int32_t n_cols, const int64_t* Ap, const int32_t* Aj,
                          std::vector<int64_t>& Cp, std::vector<int32_t>& Cj);

// ---- ordering (ordering.cpp)
void amd_order(int32_t n, const int64_t* colptr, const int32_t* rowval, std::vector<int32_t>& perm,
               double dense_factor = 10.0);
// nested dissection with AMD leaves (dissection.cpp)
void nd_order(int32_t n, const int64_t* colptr, const int32_t* rowval, std::vector<int32_t>& perm,
              int32_t leaf_size, double max_sep_frac);

struct SymbolicOptions {
    int32_t ordering = 2;          // 0 = AMD, 1 = natural (identity), 2 = nested dissection + AMD leaves
    int32_t nd_leaf = 0;           // dissection stops below this many vertices (0 = automatic)
    double nd_max_sep = 0.2;       // give a region to AMD when its separator exceeds this fraction
    int32_t max_width = 64;        // cap on supernode width (columns)
    int32_t relax_small = 8;       // always merge a last child if the merged width <= this
    double relax_zeros[3] = {0.5, 0.2, 0.05};   // allowed zero fraction for merged width <= 16, <= 48, larger
    int64_t block_panel = 12288;   // panels up to this many doubles (and at most block_below rows under the diagonal
    int32_t block_below = 256;     //   block) are factorised by ONE thread block in shared memory; larger ones are "big"
    double dense_factor = 10.0;
};

// Everything the numeric phase needs, for a symmetric matrix given by the
// lower triangle in CSC (sorted, duplicate-free, diagonal present or not).
struct Symbolic {
    int32_t n = 0;
    int64_t nnz_lower = 0;

    std::vector<int32_t> perm;      // perm[new] = old
    std::vector<int32_t> iperm;     // iperm[old] = new
    std::vector<int32_t> parent;    // column elimination tree in the new numbering (-1 root)
    std::vector<int32_t> colcount;  // nnz of column j of L incl. diagonal (exact, before relaxation)

    int32_t nsuper = 0;
    std::vector<int32_t> sn_first;   // [nsuper+1] first column of supernode s
    std::vector<int32_t> sn_of;      // [n] supernode of column j
    std::vector<int32_t> sn_parent;  // [nsuper] supernodal etree
    std::vector<int64_t> sn_rowptr;  // [nsuper+1] into sn_rows
    std::vector<int32_t> sn_rows;    // row structure (new numbering), first w entries are the supernode's own columns
    std::vector<int64_t> sn_xoff;    // [nsuper+1] offset of the h x w column-major panel in Lx

    std::vector<int64_t> asm_map;    // [nnz_lower] CSC nz -> slot in Lx

    // scatter tables of the right-looking update, indexed like sn_rows (entries of a supernode's own
    // columns are unused).  For below-row jj of supernode s, owned as a column by supernode t:
    //   row_colbase[jj] = slot of (row 0 of t's panel, column rows[jj]) in Lx
    //   row_relbase[jj] >= 0: position of rows[i] (i >= jj) in t's row list is rel[row_relbase[jj] + i]
    //                   <  0: it is i + c with c = -(row_relbase[jj] + 1) - 2^40   (contiguous match)
    //                   == kRelSearch: not tabulated (very tall supernode), binary search at run time
    std::vector<int64_t> row_colbase, row_relbase;
    std::vector<int32_t> rel;
    static constexpr int64_t kRelShift = int64_t(1) << 40;
    static constexpr int64_t kRelSearch = INT64_MIN;

    // schedule: supernodes grouped by height in the supernodal elimination tree (0 = leaf); all supernodes of
    // one height are independent, so a height is one phase of the persistent factor / solve kernels.
    //   kind 0  "small": one block (or one warp) factorises it in shared memory, multifrontal style: its
    //           contribution block C (packed lower triangle, (h-w)(h-w+1)/2 doubles at sn_coff) is pulled by the parent
    //   kind 1  small, but the parent is big: it scatters C into the ancestors' panels itself (FP64 atomics)
    //   kind 2  "big": multi-block right-looking kernels (diag / TRSM / DMMA update), pushes into the panels
    std::vector<int32_t> sn_level;       // [nsuper] height
    std::vector<uint8_t> sn_kind;        // [nsuper]
    int32_t n_levels = 0;                // number of heights
    std::vector<int32_t> child_ptr;      // [nsuper+1] into child_idx
    std::vector<int32_t> child_idx;      // children of every supernode, ascending
    std::vector<int32_t> par_pos;        // indexed like sn_rows: position of below-row i inside the PARENT's row list
    std::vector<int64_t> sn_coff;        // [nsuper+1] offsets of the contribution blocks (kind 0 / 1 only)
    // A small supernode is a task of one of four classes: a group of 256, 128, 64 or 32 threads works on it with a
    // proportional slice of the block's shared memory (class = the smallest group whose slice holds the panel).
    std::vector<int32_t> task_sn;        // small supernodes ordered by (height, class 256 first, heaviest first)
    std::vector<int32_t> phase_ptr;      // [n_levels+1] into task_sn
    std::vector<int32_t> phase_cls;      // [n_levels][5] first task of class 256, 128, 64, 32 and the end of the phase
    std::vector<int32_t> phase_nblock;   // [n_levels] tasks of a phase that are not warp (class 32) tasks
    int32_t tree_slice[4] = {0, 0, 0, 0};   // doubles of shared memory of one group, per class
    int32_t tree_cap[4] = {0, 0, 0, 0};     // ... of which panel
    int32_t tree_maxh[4] = {0, 0, 0, 0};    // ... and rows (vector scratch)
    int32_t tree_fixed = 0;                 // ... and the rest (max_w doubles, child records, ticket)
    std::vector<int32_t> big_ptr;        // [n_levels+1] into big_sn
    std::vector<int32_t> big_sn;         // big supernodes grouped by height
    int32_t n_block_tasks = 0, n_warp_tasks = 0, n_big = 0;
    int32_t small_max_panel = 0, small_max_h = 0;   // over kind 0 / 1 (shared-memory sizing of the tree kernels)

    int64_t nnz_L = 0;        // exact nnz(L) = sum colcount
    int64_t nnz_stored = 0;   // panel storage incl. relaxation zeros and the unused upper triangles
    double flops = 0;         // sum_j colcount_j^2  
    int32_t max_h = 0, max_w = 0;
};

// lower CSC (0-based): colptr[n+1], rowval sorted ascending in each column, rows >= col.
// `delayed` (optional, n flags): nodes that must not be eliminated before at least one of their
// neighbours.  For the K2 system these are the constraint rows: eliminated first, their pivot
// would be the bare dual regularisation delta_c (-1e-8, or 0 with the reference's default), which
// makes LDL' without pivoting inaccurate or impossible; eliminated right after a neighbouring
// variable x_j the pivot is -delta_c - a_ij^2 / pivot(x_j).
void analyze(int32_t n, const int64_t* colptr, const int32_t* rowval, const SymbolicOptions& opt, Symbolic& out,
             const uint8_t* delayed = nullptr);

// ---- K2 / normal-equations KKT structures -------------------------------------------------
struct KktPattern {
    int32_t kind = 0;   // 0 = K2 augmented, 1 = normal equations
    int32_t nx = 0, n_slack = 0, n_tot = 0, m = 0;
    int64_t nnzh = 0, nnzj = 0;
    int32_t N = 0;      // order of the matrix handed to the linear solver

    // Jacobian incl. slack columns, COO order [jac | slack] (normalkkt.jl:67-77), then CSR by S1
    std::vector<int32_t> jac_I, jac_J;        // 0-based COO, length nnzj + n_slack
    std::vector<int64_t> A_rowptr;            // CSR of A (m x n_tot)
    std::vector<int32_t> A_colind;
    std::vector<int64_t> A_csr_order;         // CSR slot -> COO index (0-based A_csr_map)
    std::vector<int64_t> AT_rowptr;           // CSR of A' (n_tot x m)
    std::vector<int32_t> AT_colind;
    std::vector<int64_t> AT_src;              // A' slot -> slot in A's CSR value array

    // Hessian lower triangle COO (forced lower), and full symmetric CSR for Qx
    std::vector<int32_t> hess_I, hess_J;      // 0-based
    std::vector<int64_t> Q_rowptr;            // full symmetric CSR (n_tot x n_tot)
    std::vector<int32_t> Q_colind;
    std::vector<int64_t> Q_src;               // CSR slot -> index into the hess COO values

    // matrix handed to the linear solver: lower CSC
    std::vector<int64_t> kkt_colptr;
    std::vector<int32_t> kkt_rowval;
    // K2: aug COO (block order pr_diag | hess | jac | slack | du_diag) -> CSC nz
    std::vector<int64_t> aug_csc_map;         // length n_tot + nnzh + nnzj + n_slack + m
    std::vector<int64_t> diag_slot;           // K2: [N] CSC nz of the (i, i) entry
    // normal equations: product lists  C[c] = sum_t  A[pa[t]] * D[pk[t]] * A[pb[t]]
    std::vector<int64_t> prod_ptr;            // [nnzC+1]
    std::vector<int32_t> prod_a, prod_b, prod_k;
};

void build_k2_pattern(int32_t nx, int32_t m, int64_t nnzh, const int32_t* H_I, const int32_t* H_J,
                      int64_t nnzj, const int32_t* J_I, const int32_t* J_J,
                      int32_t n_slack, const int64_t* ind_ineq, KktPattern& out);   // all indices 1-based

void build_normal_kkt_pattern(int32_t nx, int32_t m, int64_t nnzj, const int32_t* J_I, const int32_t* J_J,
                              int32_t n_slack, const int64_t* ind_ineq, KktPattern& out);   // 1-based

}  // namespace madipm
