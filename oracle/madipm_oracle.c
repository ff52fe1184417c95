/*
 * madipm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's algorithms on the per-iteration KKT hot
 * path, used only as the parity checker (tests/, __graft_entry__.smoke()) and as
 * the timed CPU baseline (bench.py cpu_baseline / --impl reference).  Nothing
 * under madqp.jl_b200/ may import, link or call this file.
 *
 * PARITY PINNING.  The reference (Julia) cannot run in this environment and
 * ships no golden vectors.  The functions restated from
 * in-repo source (coo_to_csr, build_normal_system, assemble_normal_system!,
 * transfer!) are pinned by hand-derived fixtures in tests/golden/ and by the
 * reference's one known-answer instance (simple_lp, objective 1.0).  The
 * LDL' routines restate the published algorithm of the un-vendored dependency
 * LDLFactorizations.jl 0.10.1 (T. A. Davis, "Algorithm 849: a concise sparse
 * Cholesky factorization package", ACM TOMS 31(4), 2005): parity unpinned,
 * validated against dense numpy/scipy solves only.
 *
 * All routines use the reference's 1-based indices where they restate in-repo
 * Julia, so that the loops can be compared line by line.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * coo_to_csr -- /root/reference/src/utils.jl:148-191.  1-based Ai, Aj, Bp, Bj.
 * Stable counting sort by row; keeps input order inside a row; no duplicate merge.
 * ---------------------------------------------------------------------------------------- */
void orc_coo_to_csr(int64_t n_rows, int64_t n_cols, int64_t nnz,
                    const int32_t* Ai, const int32_t* Aj, const double* Ax,
                    int32_t* Bp, int32_t* Bj, double* Bx) {
    (void)n_cols;
    for (int64_t i = 0; i <= n_rows; ++i) Bp[i] = 0;
    for (int64_t n = 0; n < nnz; ++n) Bp[Ai[n] - 1] += 1;               /* l.162-164 */
    int32_t cumsum = 1;                                                  /* l.167-173 */
    for (int64_t i = 0; i < n_rows; ++i) {
        int32_t tmp = Bp[i];
        Bp[i] = cumsum;
        cumsum += tmp;
    }
    Bp[n_rows] = (int32_t)nnz + 1;
    for (int64_t n = 0; n < nnz; ++n) {                                  /* l.175-181 */
        int32_t i = Ai[n] - 1;
        int32_t dest = Bp[i] - 1;
        Bj[dest] = Aj[n];
        Bx[dest] = Ax[n];
        Bp[i] += 1;
    }
    int32_t last = 1;                                                    /* l.183-188 */
    for (int64_t i = 0; i <= n_rows; ++i) {
        int32_t tmp = Bp[i];
        Bp[i] = last;
        last = tmp;
    }
}

/* ------------------------------------------------------------------------------------------
 * build_normal_system -- /root/reference/src/utils.jl:199-264, the reference's own
 * O(m * nnz) scan, unchanged.  Call with Cj == NULL to get the count (returned) and Cp;
 * call again with Cj allocated to fill it.  1-based.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_build_normal_system(int64_t n_rows, int64_t n_cols, const int32_t* Jtp, const int32_t* Jtj,
                                int32_t* Cp, int32_t* Cj) {
    uint8_t* xb = (uint8_t*)calloc((size_t)(n_cols > 0 ? n_cols : 1), 1);
    int64_t nnz = 0;
    if (!Cj) {
        for (int64_t i = 0; i <= n_rows; ++i) Cp[i] = 0;
        for (int64_t i = 1; i <= n_rows; ++i) {
            for (int32_t c = Jtp[i - 1]; c <= Jtp[i] - 1; ++c) xb[Jtj[c - 1] - 1] = 1;
            for (int64_t j = i; j <= n_rows; ++j)
                for (int32_t c = Jtp[j - 1]; c <= Jtp[j] - 1; ++c)
                    if (xb[Jtj[c - 1] - 1] == 1) { nnz += 1; Cp[i - 1] += 1; break; }
            for (int32_t c = Jtp[i - 1]; c <= Jtp[i] - 1; ++c) xb[Jtj[c - 1] - 1] = 0;
        }
        int32_t cumsum = 1;
        for (int64_t i = 0; i < n_rows; ++i) { int32_t tmp = Cp[i]; Cp[i] = cumsum; cumsum += tmp; }
        Cp[n_rows] = (int32_t)nnz + 1;
    } else {
        int64_t cnt = 0;
        for (int64_t i = 1; i <= n_rows; ++i) {
            for (int32_t c = Jtp[i - 1]; c <= Jtp[i] - 1; ++c) xb[Jtj[c - 1] - 1] = 1;
            for (int64_t j = i; j <= n_rows; ++j)
                for (int32_t c = Jtp[j - 1]; c <= Jtp[j] - 1; ++c)
                    if (xb[Jtj[c - 1] - 1] == 1) { Cj[cnt++] = (int32_t)j; break; }
            for (int32_t c = Jtp[i - 1]; c <= Jtp[i] - 1; ++c) xb[Jtj[c - 1] - 1] = 0;
        }
        nnz = cnt;
    }
    free(xb);
    return nnz;
}

/* ------------------------------------------------------------------------------------------
 * assemble_normal_system! -- /root/reference/src/utils.jl:266-298.  1-based.
 * Not contracted into FMAs: Julia evaluates  Cx[c] += buffer[k] * Jtx[d]  as a rounded
 * product followed by a rounded sum (compile this file with -ffp-contract=off).
 * ---------------------------------------------------------------------------------------- */
void orc_assemble_normal_system(int64_t n_rows, int64_t n_cols, const int32_t* Jtp, const int32_t* Jtj,
                                const double* Jtx, const int32_t* Cp, const int32_t* Cj, double* Cx,
                                const double* Dx) {
    double* buffer = (double*)calloc((size_t)(n_cols > 0 ? n_cols : 1), sizeof(double));
    for (int64_t i = 1; i <= n_rows; ++i) {
        for (int32_t c = Jtp[i - 1]; c <= Jtp[i] - 1; ++c) {
            int32_t j = Jtj[c - 1];
            buffer[j - 1] = Jtx[c - 1] * Dx[j - 1];
        }
        for (int32_t c = Cp[i - 1]; c <= Cp[i] - 1; ++c) {
            int32_t j = Cj[c - 1];
            double acc = 0.0;
            for (int32_t d = Jtp[j - 1]; d <= Jtp[j] - 1; ++d) {
                int32_t k = Jtj[d - 1];
                double prod = buffer[k - 1] * Jtx[d - 1];
                acc = acc + prod;
            }
            Cx[c - 1] = acc;
        }
        for (int32_t c = Jtp[i - 1]; c <= Jtp[i] - 1; ++c) buffer[Jtj[c - 1] - 1] = 0.0;
    }
    free(buffer);
}

/* ------------------------------------------------------------------------------------------
 * transfer! -- K2 build_kkt!, semantics stated in /root/reference/scripts/cuda_wrapper.jl:17-28:
 * dest .= 0; dest[map[k]] += src[k], serially in k.  map is 1-based.
 * ---------------------------------------------------------------------------------------- */
void orc_transfer(int64_t ndest, double* dest, int64_t nsrc, const double* src, const int64_t* map) {
    for (int64_t i = 0; i < ndest; ++i) dest[i] = 0.0;
    for (int64_t k = 0; k < nsrc; ++k) dest[map[k] - 1] += src[k];
}

/* ------------------------------------------------------------------------------------------
 * Sparse LDL' (Davis, Algorithm 849), as used by LDLFactorizations.jl (un-vendored; call sites
 * /root/reference/src/utils.jl:57-59, scripts/benchmarks_cpu.jl:36).  0-based.
 * Input: upper triangle of the ALREADY PERMUTED matrix in CSC (column k holds rows i <= k).
 * ---------------------------------------------------------------------------------------- */
void orc_ldl_symbolic(int64_t n, const int64_t* Cp, const int32_t* Ci, int64_t* Lp, int32_t* parent,
                      int32_t* Lnz, int32_t* flag) {
    for (int64_t k = 0; k < n; ++k) {
        parent[k] = -1;
        flag[k] = (int32_t)k;
        Lnz[k] = 0;
        for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
            int32_t i = Ci[p];
            if (i < k) {
                for (; flag[i] != k; i = parent[i]) {
                    if (parent[i] == -1) parent[i] = (int32_t)k;
                    Lnz[i]++;
                    flag[i] = (int32_t)k;
                }
            }
        }
    }
    Lp[0] = 0;
    for (int64_t k = 0; k < n; ++k) Lp[k + 1] = Lp[k] + Lnz[k];
}

/* returns n on success, or the column k (0-based) whose pivot is exactly zero */
int64_t orc_ldl_numeric(int64_t n, const int64_t* Cp, const int32_t* Ci, const double* Cx,
                        const int64_t* Lp, const int32_t* parent, int32_t* Lnz, int32_t* Li, double* Lx,
                        double* D, double* Y, int32_t* pattern, int32_t* flag) {
    for (int64_t k = 0; k < n; ++k) {
        Y[k] = 0.0;
        int64_t top = n;
        flag[k] = (int32_t)k;
        Lnz[k] = 0;
        for (int64_t p = Cp[k]; p < Cp[k + 1]; ++p) {
            int32_t i = Ci[p];
            if (i <= k) {
                Y[i] += Cx[p];
                int64_t len = 0;
                for (; flag[i] != k; i = parent[i]) {
                    pattern[len++] = i;
                    flag[i] = (int32_t)k;
                }
                while (len > 0) pattern[--top] = pattern[--len];
            }
        }
        D[k] = Y[k];
        Y[k] = 0.0;
        for (; top < n; ++top) {
            int32_t i = pattern[top];
            double yi = Y[i];
            Y[i] = 0.0;
            int64_t p2 = Lp[i] + Lnz[i];
            for (int64_t p = Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
            double l_ki = yi / D[i];
            D[k] -= l_ki * yi;
            Li[p2] = (int32_t)k;
            Lx[p2] = l_ki;
            Lnz[i]++;
        }
        if (D[k] == 0.0) return k;
    }
    return n;
}

void orc_ldl_solve(int64_t n, double* x, const int64_t* Lp, const int32_t* Li, const double* Lx, const double* D) {
    for (int64_t j = 0; j < n; ++j) {
        double xj = x[j];
        for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) x[Li[p]] -= Lx[p] * xj;
    }
    for (int64_t j = 0; j < n; ++j) x[j] /= D[j];
    for (int64_t j = n - 1; j >= 0; --j) {
        double xj = x[j];
        for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) xj -= Lx[p] * x[Li[p]];
        x[j] = xj;
    }
}

/* CSR y = alpha * A x + beta * y (0-based), for the timed CPU baseline */
void orc_csr_matvec(int64_t n_rows, const int64_t* Ap, const int32_t* Aj, const double* Ax,
                    const double* x, double* y, double alpha, double beta) {
    for (int64_t i = 0; i < n_rows; ++i) {
        double acc = 0.0;
        for (int64_t p = Ap[i]; p < Ap[i + 1]; ++p) acc += Ax[p] * x[Aj[p]];
        y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * acc;
    }
}
