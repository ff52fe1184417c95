"""oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy + C (oracle/madipm_oracle.c) restatement of the reference's MPC solver
on the CPU, function by function, for use as the parity checker in tests/,
in __graft_entry__.smoke() and as the timed CPU baseline of bench.py.
Nothing under madqp.jl_b200/ may import this module.

Every function cites the reference lines it follows.  Parts that live in the
un-vendored dependency MadNLP 0.8.7 are marked [RECALLED]: they
are restated from the algebra the in-repo code forces (SURVEY.md 8a note A)
and parity for them is unpinned.

PARITY PINNING: see the header of madipm_oracle.c.  The only reference-held
known answer is simple_lp() (objective 1.0, /root/reference/test/runtests.jl:24-55).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SOLVE_SUCCEEDED = "SOLVE_SUCCEEDED"
MAXIMUM_ITERATIONS_EXCEEDED = "MAXIMUM_ITERATIONS_EXCEEDED"
ERROR_IN_STEP_COMPUTATION = "ERROR_IN_STEP_COMPUTATION"
INTERNAL_ERROR = "INTERNAL_ERROR"


class SolveException(Exception):
    """MadNLP.SolveException (/root/reference/src/linear_solver.jl:41-43)."""


# -O3 with AVX2 code generation (x86-64-v3 rather than -march=native: the library is built in one container and runs
# on another box); no FMA contraction and no -ffast-math, so the arithmetic is the reference's operation by operation.
C_FLAGS = "gcc -O3 -march=x86-64-v3 -ffp-contract=off"


def build(force=False):
    """Compile the C restatement into oracle/liboracle.so (gcc, no FMA contraction)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "madipm_oracle.c")
    stamp = so + ".flags"
    old = open(stamp).read() if os.path.exists(stamp) else ""
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src) or old != C_FLAGS:
        subprocess.check_call(C_FLAGS.split() + ["-fPIC", "-shared", "-o", so, src, "-lm"])
        with open(stamp, "w") as f:
            f.write(C_FLAGS)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_build_normal_system.restype = ctypes.c_int64
        _LIB.orc_ldl_numeric.restype = ctypes.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


_i32, _i64, _f64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_double


# ---------------------------------------------------------------------------------------------
# sparse utilities (1-based, as in the reference)
# ---------------------------------------------------------------------------------------------
def coo_to_csr(n_rows, n_cols, Ai, Aj, Ax):
    """/root/reference/src/utils.jl:148-191.  1-based Int32 in and out."""
    Ai = np.ascontiguousarray(Ai, dtype=np.int32)
    Aj = np.ascontiguousarray(Aj, dtype=np.int32)
    Ax = np.ascontiguousarray(Ax, dtype=np.float64)
    nnz = Ai.size
    Bp = np.zeros(n_rows + 1, dtype=np.int32)
    Bj = np.zeros(nnz, dtype=np.int32)
    Bx = np.zeros(nnz, dtype=np.float64)
    lib().orc_coo_to_csr(_i64(n_rows), _i64(n_cols), _i64(nnz), _p(Ai, _i32), _p(Aj, _i32), _p(Ax, _f64),
                         _p(Bp, _i32), _p(Bj, _i32), _p(Bx, _f64))
    return Bp, Bj, Bx


def build_normal_system(n_rows, n_cols, Jtp, Jtj):
    """/root/reference/src/utils.jl:199-264 (the reference's own O(m nnz) scan)."""
    Jtp = np.ascontiguousarray(Jtp, dtype=np.int32)
    Jtj = np.ascontiguousarray(Jtj, dtype=np.int32)
    Cp = np.zeros(n_rows + 1, dtype=np.int32)
    nnz = lib().orc_build_normal_system(_i64(n_rows), _i64(n_cols), _p(Jtp, _i32), _p(Jtj, _i32), _p(Cp, _i32), None)
    Cj = np.zeros(max(nnz, 1), dtype=np.int32)
    lib().orc_build_normal_system(_i64(n_rows), _i64(n_cols), _p(Jtp, _i32), _p(Jtj, _i32), _p(Cp, _i32), _p(Cj, _i32))
    return Cp, Cj[:nnz]


def build_normal_system_fast(n_rows, n_cols, Jtp, Jtj):
    """Same (Cp, Cj) as build_normal_system via scipy's boolean A A' (used for the
    large CPU-baseline runs where the O(m nnz) scan is infeasible)."""
    nnz = int(Jtp[-1] - 1)
    A = sp.csr_matrix((np.ones(nnz), np.asarray(Jtj[:nnz], dtype=np.int64) - 1, np.asarray(Jtp, dtype=np.int64) - 1),
                      shape=(n_rows, n_cols))
    C = sp.tril((A @ A.T).tocsc(), format="csc")
    C.sort_indices()
    return (C.indptr + 1).astype(np.int32), (C.indices + 1).astype(np.int32)


def assemble_normal_system(n_rows, n_cols, Jtp, Jtj, Jtx, Cp, Cj, Cx, Dx):
    """/root/reference/src/utils.jl:266-298."""
    lib().orc_assemble_normal_system(_i64(n_rows), _i64(n_cols), _p(Jtp, _i32), _p(Jtj, _i32), _p(Jtx, _f64),
                                     _p(Cp, _i32), _p(Cj, _i32), _p(Cx, _f64), _p(Dx, _f64))


def transfer(dest, src, to_map):
    """K2 build_kkt!: /root/reference/scripts/cuda_wrapper.jl:17-28 (serial semantics)."""
    lib().orc_transfer(_i64(dest.size), _p(dest, _f64), _i64(src.size), _p(src, _f64), _p(to_map, _i64))


def coo_to_csc(m, n, I, J):
    """MadNLP.coo_to_csc [RECALLED]: CSC = sparse(I, J, V) (sorted, duplicates merged),
    map[k] = nz index of (I[k], J[k]).  1-based in and out."""
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    key = (J - 1) * m + (I - 1)
    uniq, inv = np.unique(key, return_inverse=True)
    rows = (uniq % m).astype(np.int32) + 1
    cols = uniq // m
    colptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(colptr, cols + 1, 1)
    colptr = (np.cumsum(colptr) + 1).astype(np.int32)
    return colptr, rows, (inv + 1).astype(np.int64)


# ---------------------------------------------------------------------------------------------
# LDLSolver restatement (MadNLP.LDLSolver -> LDLFactorizations.jl, un-vendored)
# ---------------------------------------------------------------------------------------------
def scale_augmented_system_coo(I, J, src_V, scaling, n, m):
    """_scale_augmented_system_coo_kernel! (/root/reference/scripts/cuda_wrapper.jl:90-111), all entries at once
    (1-based I, J; k = position in the COO arrays): k <= n copied (pr_diag), Hessian block V * s_i * s_j, Jacobian
    block V * s_j, dual-regularisation block copied.  Returns dest_V."""
    I, J = np.asarray(I), np.asarray(J)
    src_V = np.asarray(src_V, dtype=np.float64)
    k = np.arange(1, src_V.size + 1)
    dest = src_V.copy()
    hess = (k > n) & (I <= n) & (J <= n)
    dest[hess] = src_V[hess] * scaling[I[hess] - 1] * scaling[J[hess] - 1]
    jac = (k > n) & (I >= n + 1) & (I <= n + m) & (J <= n)
    dest[jac] = src_V[jac] * scaling[J[jac] - 1]
    return dest


def standard_form_qp(qp):
    """standard_form_qp, /root/reference/scripts/common.jl:156-288, loop for loop (1-based inside, like the Julia)."""
    import dataclasses
    n, m = qp.nvar, qp.ncon
    lvar, uvar, lcon, ucon = (np.asarray(a, dtype=np.float64) for a in (qp.lvar, qp.uvar, qp.lcon, qp.ucon))
    ind_ineq = [i for i in range(1, m + 1) if lcon[i - 1] < ucon[i - 1]]                      # l.164-168
    ns = len(ind_ineq)
    ind_rng, ind_only_ub, ind_fixed, xu = [], [], [], []
    for i in range(1, n + 1):                                                                # l.176-188
        if lvar[i - 1] == uvar[i - 1]:
            ind_fixed.append(i)
        elif -np.inf < lvar[i - 1] < uvar[i - 1] < np.inf:
            ind_rng.append(i); xu.append(uvar[i - 1])
        elif uvar[i - 1] < np.inf:
            ind_only_ub.append(i)
    for k, i in enumerate(ind_ineq, start=1):                                                # l.191-199
        if -np.inf < lcon[i - 1] < ucon[i - 1] < np.inf:
            ind_rng.append(k + n); xu.append(ucon[i - 1])
        elif ucon[i - 1] < np.inf:
            ind_only_ub.append(k + n)
    nw = len(ind_rng)
    nvar, ncon = n + ns + nw, m + nw
    Bi, Bj, Bx = [], [], []
    for k, i in enumerate(ind_ineq, start=1):                                                # l.214-219
        Bi.append(i); Bj.append(n + k); Bx.append(-1.0)
    for k, i in enumerate(ind_rng, start=1):                                                 # l.221-230
        Bi.append(m + k); Bj.append(i); Bx.append(1.0)
        Bi.append(m + k); Bj.append(k + n + ns); Bx.append(1.0)
    lcon_, ucon_ = np.zeros(ncon), np.zeros(ncon)
    for i in range(1, m + 1):                                                                # l.238-248
        if not (lcon[i - 1] < ucon[i - 1]):
            lcon_[i - 1], ucon_[i - 1] = lcon[i - 1], ucon[i - 1]
    for k, i in enumerate(ind_rng, start=1):                                                 # l.249-252
        lcon_[m + k - 1] = xu[k - 1]; ucon_[m + k - 1] = xu[k - 1]
    ineq0 = np.array(ind_ineq, dtype=np.int64) - 1
    lvar_ = np.concatenate([lvar, lcon[ineq0], np.zeros(nw)])                                # l.254-259
    uvar_ = np.concatenate([uvar, ucon[ineq0], np.full(nw, np.inf)])
    for i in ind_rng:
        uvar_[i - 1] = np.inf
    for i in ind_fixed:
        uvar_[i - 1] = uvar[i - 1]
    z = np.zeros(ns + nw)
    return dataclasses.replace(
        qp, name=qp.name + "-std", nvar=nvar, ncon=ncon, c=np.concatenate([qp.c, z]),
        A_rows=np.concatenate([qp.A_rows, np.array(Bi, dtype=np.int64) - 1]).astype(qp.A_rows.dtype),
        A_cols=np.concatenate([qp.A_cols, np.array(Bj, dtype=np.int64) - 1]).astype(qp.A_cols.dtype),
        A_vals=np.concatenate([np.asarray(qp.A_vals, dtype=np.float64), np.array(Bx, dtype=np.float64)]),
        lcon=lcon_, ucon=ucon_, lvar=lvar_, uvar=uvar_, x0=np.concatenate([qp.x0, z]),
        y0=np.concatenate([qp.y0, np.zeros(nw)]), meta=dict(qp.meta, standard_form=True))


def scale_qp(qp, max_iter=10, tol=0.0):
    """scale_qp, /root/reference/scripts/common.jl:46-100.  (Dr, Dc) = HSL.mc77(A, 0) there: HSL is proprietary and
    un-vendored, so the published algorithm is restated (D. Ruiz, RAL-TR-2001-034: simultaneous row / column sweeps
    D <- D * sqrt(inf-norm of the currently scaled row / column); [RECALLED] MC77 defaults: 10 sweeps, no convergence
    test).  Parity with MC77's exact factors is unpinned; the rescaling formulas are the reference's (l.38-44, 70-97).
    Returns (scaled copy of qp, Dr, Dc, sweeps, err)."""
    import dataclasses
    m, n = qp.ncon, qp.nvar
    I, J, V = qp.A_rows, qp.A_cols, np.asarray(qp.A_vals, dtype=np.float64)
    Dr, Dc = np.ones(m), np.ones(n)
    sweeps, err = 0, 0.0
    for _ in range(max_iter):
        v = np.abs(V) / (Dr[I] * Dc[J])
        rmax, cmax = np.zeros(m), np.zeros(n)
        np.maximum.at(rmax, I, v)
        np.maximum.at(cmax, J, v)
        err = 0.0
        for D, mx in ((Dr, rmax), (Dc, cmax)):
            nz = mx > 0.0
            D[nz] = D[nz] * np.sqrt(mx[nz])
            if nz.any():
                err = max(err, float(np.max(np.abs(1.0 - mx[nz]))))
        sweeps += 1
        if tol > 0.0 and err <= tol:
            break
    out = dataclasses.replace(
        qp, name=qp.name + "-ruiz", A_vals=V / (Dr[I] * Dc[J]),                       # _scale_coo!(As, Dr, Dc)
        H_vals=np.asarray(qp.H_vals, dtype=np.float64) / (Dc[qp.H_rows] * Dc[qp.H_cols]),   # _scale_coo!(Hs, Dc, Dc)
        c=qp.c / Dc, lvar=qp.lvar * Dc, uvar=qp.uvar * Dc, lcon=qp.lcon / Dr, ucon=qp.ucon / Dr,
        x0=qp.x0 * Dc, y0=qp.y0 / Dr, meta=dict(qp.meta, ruiz=True))
    return out, Dr, Dc, sweeps, err


class LDLSolver:
    """Scalar up-looking LDL' on a fixed fill-reducing permutation, no pivoting.

    csc: lower triangle (colptr, rowval, nzval view) 1-based, as handed to the
    linear-solver constructor at /root/reference/src/KKT/normalkkt.jl:99-101.
    perm: 0-based permutation, perm[new] = old (the reference gets it from AMD.jl;
    here it is an input because that ordering cannot be reproduced -- SURVEY 7).
    """

    def __init__(self, n, colptr, rowval, nzval, perm=None):
        self.n = n
        self.nzval = nzval
        perm = np.arange(n, dtype=np.int64) if perm is None else np.asarray(perm, dtype=np.int64)
        assert np.array_equal(np.sort(perm), np.arange(n)), "perm is not a permutation"
        self.perm = perm
        iperm = np.empty(n, dtype=np.int64)
        iperm[perm] = np.arange(n)
        cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptr.astype(np.int64)))
        rows = rowval.astype(np.int64) - 1
        a, b = iperm[rows], iperm[cols]
        r, c = np.minimum(a, b), np.maximum(a, b)          # upper triangle of P A P'
        order = np.lexsort((r, c))
        self._src = order                                  # upper CSC slot -> lower CSC nz
        self.Ci = r[order].astype(np.int32)
        self.Cp = np.zeros(n + 1, dtype=np.int64)
        np.add.at(self.Cp, c + 1, 1)
        self.Cp = np.cumsum(self.Cp)
        self.Lp = np.zeros(n + 1, dtype=np.int64)
        self.parent = np.zeros(n, dtype=np.int32)
        self.Lnz = np.zeros(n, dtype=np.int32)
        self._flag = np.zeros(n, dtype=np.int32)
        lib().orc_ldl_symbolic(_i64(n), _p(self.Cp, _i64), _p(self.Ci, _i32), _p(self.Lp, _i64),
                               _p(self.parent, _i32), _p(self.Lnz, _i32), _p(self._flag, _i32))
        self.colcount = self.Lnz.copy() + 1
        nnzL = int(self.Lp[-1])
        self.Li = np.zeros(max(nnzL, 1), dtype=np.int32)
        self.Lx = np.zeros(max(nnzL, 1), dtype=np.float64)
        self.D = np.zeros(n)
        self._Y = np.zeros(n)
        self._pattern = np.zeros(n, dtype=np.int32)
        self.factorized = False
        self.nnz_L = nnzL + n
        self.flops = float(np.sum(self.colcount.astype(np.float64) ** 2))

    def factorize(self):
        Cx = np.ascontiguousarray(self.nzval[self._src])
        k = lib().orc_ldl_numeric(_i64(self.n), _p(self.Cp, _i64), _p(self.Ci, _i32), _p(Cx, _f64),
                                  _p(self.Lp, _i64), _p(self.parent, _i32), _p(self.Lnz, _i32), _p(self.Li, _i32),
                                  _p(self.Lx, _f64), _p(self.D, _f64), _p(self._Y, _f64), _p(self._pattern, _i32),
                                  _p(self._flag, _i32))
        self.factorized = (k == self.n) and bool(np.all(np.isfinite(self.D)))
        return self.factorized

    def solve(self, rhs):
        y = np.ascontiguousarray(rhs[self.perm])
        lib().orc_ldl_solve(_i64(self.n), _p(y, _f64), _p(self.Lp, _i64), _p(self.Li, _i32), _p(self.Lx, _f64),
                            _p(self.D, _f64))
        rhs[self.perm] = y
        return rhs

    def inertia(self):
        return int((self.D > 0).sum()), int((self.D == 0).sum()), int((self.D < 0).sum())


# ---------------------------------------------------------------------------------------------
# problem plumbing [RECALLED MadNLP]: index sets, initial point, scaling
# ---------------------------------------------------------------------------------------------
@dataclass
class IndexConstraints:
    ind_eq: np.ndarray
    ind_ineq: np.ndarray
    ind_lb: np.ndarray
    ind_ub: np.ndarray
    ind_llb: np.ndarray
    ind_uub: np.ndarray


def make_parameter(qp):
    """fixed_variable_treatment = MakeParameter (/root/reference/src/utils.jl:81, the default for the K2 systems;
    implemented in MadNLP's SparseCallback, un-vendored, [RECALLED]): a variable with lvar == uvar is a parameter of the
    model, not an unknown -- its direction is identically zero, the entries of its Jacobian column and Hessian row / column
    leave the KKT system, and its value enters the constraints and the gradient as a constant.  Restated as the equivalent
    model in the free variables only, loop for loop over the COO entries:
        c_free += H[free, fixed] x_fixed,  c0 += c_fixed' x_fixed + 1/2 x_fixed' H[fixed, fixed] x_fixed,
        (lcon, ucon) -= A[:, fixed] x_fixed.
    Returns (reduced model, fixed indices, free indices, x_fixed), 0-based; (qp, empty, all, empty) without fixed variables."""
    import dataclasses
    n = qp.nvar
    lvar, uvar = np.asarray(qp.lvar, dtype=np.float64), np.asarray(qp.uvar, dtype=np.float64)
    fixed = [i for i in range(n) if lvar[i] == uvar[i]]
    free = [i for i in range(n) if not lvar[i] == uvar[i]]
    if not fixed:
        return qp, np.zeros(0, dtype=np.int64), np.arange(n), np.zeros(0)
    new_index = {i: k for k, i in enumerate(free)}
    is_fixed = set(fixed)
    xf = {i: lvar[i] for i in fixed}
    c = np.asarray(qp.c, dtype=np.float64).copy()
    c0 = float(qp.c0)
    for i in fixed:
        c0 += c[i] * xf[i]
    Hr, Hc, Hv = [], [], []
    for r, cc, v in zip(qp.H_rows, qp.H_cols, qp.H_vals):              # lower triangle of a symmetric matrix
        r, cc, v = int(r), int(cc), float(v)
        if r in is_fixed and cc in is_fixed:
            c0 += (0.5 if r == cc else 1.0) * v * xf[r] * xf[cc]
        elif r in is_fixed:
            c[cc] += v * xf[r]
        elif cc in is_fixed:
            c[r] += v * xf[cc]
        else:
            Hr.append(new_index[r]); Hc.append(new_index[cc]); Hv.append(v)
    shift = np.zeros(qp.ncon)
    Ar, Ac, Av = [], [], []
    for r, cc, v in zip(qp.A_rows, qp.A_cols, qp.A_vals):
        r, cc, v = int(r), int(cc), float(v)
        if cc in is_fixed:
            shift[r] += v * xf[cc]
        else:
            Ar.append(r); Ac.append(new_index[cc]); Av.append(v)
    idx = lambda a, like: np.array(a, dtype=np.asarray(like).dtype)
    free_a = np.array(free, dtype=np.int64)
    red = dataclasses.replace(
        qp, name=qp.name + "-free", nvar=len(free), c=c[free_a], c0=c0, H_rows=idx(Hr, qp.H_rows), H_cols=idx(Hc, qp.H_cols),
        H_vals=np.array(Hv, dtype=np.float64), A_rows=idx(Ar, qp.A_rows), A_cols=idx(Ac, qp.A_cols),
        A_vals=np.array(Av, dtype=np.float64), lcon=np.asarray(qp.lcon, dtype=np.float64) - shift,
        ucon=np.asarray(qp.ucon, dtype=np.float64) - shift, lvar=lvar[free_a], uvar=uvar[free_a],
        x0=np.asarray(qp.x0, dtype=np.float64)[free_a], meta=dict(qp.meta, make_parameter=len(fixed)))
    return red, np.array(fixed, dtype=np.int64), free_a, np.array([xf[i] for i in fixed])


def get_index_constraints(lvar, uvar, lcon, ucon):
    """MadNLP.get_index_constraints [RECALLED]; 0-based results.  Call site
    /root/reference/src/structure.jl:95-102.  (Fixed variables never get here: make_parameter removed them.)"""
    if np.any(lvar == uvar):
        raise NotImplementedError("fixed variables reach get_index_constraints: apply make_parameter first")
    ind_eq = np.nonzero(lcon == ucon)[0]
    ind_ineq = np.nonzero(lcon != ucon)[0]
    xl = np.concatenate([lvar, lcon[ind_ineq]])
    xu = np.concatenate([uvar, ucon[ind_ineq]])
    lb, ub = xl > -np.inf, xu < np.inf
    return IndexConstraints(ind_eq, ind_ineq, np.nonzero(lb)[0], np.nonzero(ub)[0],
                            np.nonzero(lb & ~ub)[0], np.nonzero(~lb & ub)[0])


def _initialize_variable(x, l, u, bound_push, bound_fac):
    """MadNLP.initialize_variables! [RECALLED] (Ipopt's push into the interior)."""
    out = x.copy()
    both = np.isfinite(l) & np.isfinite(u)
    lo = np.isfinite(l) & ~np.isfinite(u)
    up = ~np.isfinite(l) & np.isfinite(u)
    with np.errstate(invalid="ignore"):
        pl = np.minimum(bound_push * np.maximum(1.0, np.abs(l)), bound_fac * (u - l))
        pu = np.minimum(bound_push * np.maximum(1.0, np.abs(u)), bound_fac * (u - l))
        out[both] = np.minimum(u - pu, np.maximum(l + pl, x))[both]
        out[lo] = np.maximum(l + bound_push * np.maximum(1.0, np.abs(l)), x)[lo]
        out[up] = np.minimum(u - bound_push * np.maximum(1.0, np.abs(u)), x)[up]
    return out


@dataclass
class Options:
    """IPMOptions defaults, /root/reference/src/utils.jl:69-103."""
    tol: float = 1e-8
    max_iter: int = 3000
    scaling: bool = True
    bound_push: float = 1e-2
    bound_fac: float = 1e-2
    bound_relax_factor: float = 1e-8
    regularization: tuple = ("fixed", 1e-8, 0.0)     # FixedRegularization(1e-8, 0.0)
    step_rule: tuple = ("adaptive", 0.99)             # AdaptiveStep(0.99)
    max_ncorr: int = 0
    mu_init: float = 1e-1
    mu_min: float = 1e-11
    tol_linear_solve: float = 1e-8
    check_residual: bool = False
    kkt_system: str = "K2"                            # "K2" | "NORMAL" | "K2.5" (MadNLP.ScaledSparseKKTSystem)


class MPCOracle:
    """MPCSolver restated (/root/reference/src/structure.jl:77-176, solver.jl)."""

    def __init__(self, qp, perm=None, **kw):
        self.opt = Options(**kw)
        self.qp_full = qp
        qp, self._fixed, self._free, self._x_fixed = make_parameter(qp)        # utils.jl:81 (MakeParameter)
        self.qp = qp
        opt = self.opt
        ic = get_index_constraints(qp.lvar, qp.uvar, qp.lcon, qp.ucon)
        self.ic = ic
        nx, m = qp.nvar, qp.ncon
        ns = ic.ind_ineq.size
        n = nx + ns
        self.nx, self.ns, self.n, self.m = nx, ns, n, m
        self.ind_lb, self.ind_ub = ic.ind_lb, ic.ind_ub
        self.nlb, self.nub = ic.ind_lb.size, ic.ind_ub.size
        if opt.kkt_system not in ("K2", "NORMAL", "K2.5"):
            raise ValueError(f"unknown kkt_system {opt.kkt_system!r}")
        if opt.kkt_system == "NORMAL" and qp.nnzh > 0:
            raise ValueError("NormalKKTSystem supports only linear programs")   # normalkkt.jl:43-46
        self._k2 = opt.kkt_system in ("K2", "K2.5")        # both share the augmented pattern
        self._scaled = opt.kkt_system == "K2.5"

        # ---- Jacobian incl. slack columns, COO -> CSR (normalkkt.jl:67-91)
        nnzj = qp.nnzj
        I = np.concatenate([qp.A_rows + 1, ic.ind_ineq + 1]).astype(np.int32)
        J = np.concatenate([qp.A_cols + 1, np.arange(nx + 1, n + 1)]).astype(np.int32)
        Ap, Aj, Ax = coo_to_csr(m, n, I, J, np.arange(1, nnzj + ns + 1, dtype=np.float64))
        self.A_csr_map = Ax.astype(np.int64)                 # normalkkt.jl:88
        self.Ap, self.Aj = Ap, Aj
        self.jac_V = np.zeros(nnzj + ns)                     # A_coo.V; first nnzj = jac view
        self.AT_nzval = np.zeros(nnzj + ns)                  # AT.nzval == CSR values of A

        # ---- KKT system
        self.reg = np.zeros(n)
        self.pr_diag = np.zeros(n)
        self.du_diag = np.zeros(m)
        self.l_diag = np.zeros(self.nlb)
        self.u_diag = np.zeros(self.nub)
        self.l_lower = np.zeros(self.nlb)
        self.u_lower = np.zeros(self.nub)
        self.scaling_factor = np.ones(n)                    # ScaledSparseKKTSystem.scaling_factor
        if self._k2:
            # aug COO: pr_diag | hess | jac | slack | du_diag  [RECALLED; cuda_wrapper.jl:95-107]
            hI = np.maximum(qp.H_rows, qp.H_cols) + 1
            hJ = np.minimum(qp.H_rows, qp.H_cols) + 1
            nh = qp.nnzh
            aI = np.concatenate([np.arange(1, n + 1), hI, I.astype(np.int64) + n, np.arange(n + 1, n + m + 1)])
            aJ = np.concatenate([np.arange(1, n + 1), hJ, J.astype(np.int64), np.arange(n + 1, n + m + 1)])
            self.aug_V = np.zeros(aI.size)
            self._aug_I, self._aug_J = aI, aJ               # K2.5 scales the COO values by block (cuda_wrapper.jl:90-116)
            self._off_hess, self._off_jac, self._off_du = n, n + nh, n + nh + nnzj + ns
            self.aug_colptr, self.aug_rowval, self.aug_csc_map = coo_to_csc(n + m, n + m, aI, aJ)
            self.aug_nzval = np.zeros(self.aug_rowval.size)
            self.N = n + m
            self.hess_V = np.zeros(nh)
            # symmetric H for mul! / obj
            self._H = None
        else:
            t0 = time.time()
            if m * max(nnzj, 1) <= 5_000_000:
                Cp, Cj = build_normal_system(m, n, Ap, Aj)
            else:
                Cp, Cj = build_normal_system_fast(m, n, Ap, Aj)
            self.aug_colptr, self.aug_rowval = Cp, Cj
            self.aug_nzval = np.zeros(Cj.size)
            self.N = m
            self.buffer_n = np.zeros(n)
            self.buffer_m = np.zeros(m)
            self._t_pattern = time.time() - t0
        self.linear_solver = LDLSolver(self.N, self.aug_colptr, self.aug_rowval, self.aug_nzval, perm)

        # ---- vectors (structure.jl:123-151).  UnreducedKKTVector: [x | y | zl | zu]
        self.x = np.zeros(n); self.xl = np.zeros(n); self.xu = np.zeros(n)
        self.zl = np.zeros(n); self.zu = np.zeros(n); self.f = np.zeros(n)
        self.y = np.zeros(m); self.c = np.zeros(m); self.rhs = np.zeros(m); self.jacl = np.zeros(n)
        L = n + m + self.nlb + self.nub
        self.d = np.zeros(L); self.p = np.zeros(L); self._w1 = np.zeros(L); self._w2 = np.zeros(L)
        self.correction_lb = np.zeros(self.nlb); self.correction_ub = np.zeros(self.nub)
        self.obj_val = 0.0
        self.obj_scale = 1.0
        self.con_scale = np.ones(m)
        self.mu = 0.0; self.alpha_p = 0.0; self.alpha_d = 0.0; self.del_w = 0.0; self.del_c = 0.0
        self.inf_pr = self.inf_du = self.inf_compl = 0.0
        self.k = 0
        self.status = "INITIAL"
        self.trace = []
        self.t_factor = 0.0
        self.n_factor = 0

    # ---- KKT vector views (MadNLP.primal / dual / dual_lb / dual_ub) -------------------------
    def _views(self, w):
        n, m, nlb = self.n, self.m, self.nlb
        return w[:n], w[n:n + m], w[n + m:n + m + nlb], w[n + m + nlb:]

    # ---- model callbacks [RECALLED wrappers]; QuadraticModels obj/grad/cons ------------------
    def _A(self):
        return sp.csr_matrix((self.AT_nzval, self.Aj.astype(np.int64) - 1, self.Ap.astype(np.int64) - 1),
                             shape=(self.m, self.n))

    def _Hsym(self):
        qp = self.qp
        if self._Hs is None:
            H = sp.coo_matrix((self.hess_V, (np.maximum(qp.H_rows, qp.H_cols), np.minimum(qp.H_rows, qp.H_cols))),
                              shape=(self.n, self.n)).tocsr()
            self._Hs = H + sp.tril(H, -1).T
        return self._Hs

    def eval_f(self):
        # obj = c0 + c'x + x'Hx/2  (/root/reference/scripts/qp_gpu.jl:29-33), times obj_scale
        x = self.x
        v = self._Hsym() @ x
        return (self.qp.c0 * self.obj_scale) + float(self._cost @ x) + 0.5 * float(v @ x)

    def eval_grad_f(self):
        # grad = Hx + c (/root/reference/scripts/qp_gpu.jl:35-40); slack part is zero
        self.f[:] = self._Hsym() @ self.x + self._cost

    def eval_cons(self):
        # c = con_scale .* (A x) - s - rhs   [RECALLED eval_cons_wrapper!]
        self.c[:] = self._Amat @ self.x - self.rhs

    # ---- initialize! -------------------------------------------------------------------------
    def initialize(self):
        """/root/reference/src/solver.jl:127-182."""
        opt, qp, ic = self.opt, self.qp, self.ic
        nx, ns, n, m = self.nx, self.ns, self.n, self.m
        # MadNLP.initialize!(cb, ...) [RECALLED]
        tol = opt.bound_relax_factor
        lvar = qp.lvar - np.maximum(1.0, np.abs(qp.lvar)) * tol
        uvar = qp.uvar + np.maximum(1.0, np.abs(qp.uvar)) * tol
        x0 = _initialize_variable(qp.x0.astype(float), lvar, uvar, opt.bound_push, opt.bound_fac)
        A0 = sp.csr_matrix((qp.A_vals, (qp.A_rows, qp.A_cols)), shape=(m, nx))
        cons0 = A0 @ x0
        ls = qp.lcon[ic.ind_ineq] - np.maximum(1.0, np.abs(qp.lcon[ic.ind_ineq])) * tol
        us = qp.ucon[ic.ind_ineq] + np.maximum(1.0, np.abs(qp.ucon[ic.ind_ineq])) * tol
        s0 = _initialize_variable(cons0[ic.ind_ineq], ls, us, opt.bound_push, opt.bound_fac)
        self.x[:] = np.concatenate([x0, s0])
        self.xl[:] = np.concatenate([lvar, ls])
        self.xu[:] = np.concatenate([uvar, us])
        self.y[:] = qp.y0
        self.rhs[:] = np.where(qp.lcon == qp.ucon, qp.lcon, 0.0)
        self.jacl[:] = 0.0

        # set_scaling! [RECALLED], nlp_scaling_max_gradient = 100 (solver.jl:148-159)
        self.con_scale[:] = 1.0
        self.obj_scale = 1.0
        if opt.scaling:
            rowmax = np.ones(m)
            np.maximum.at(rowmax, qp.A_rows, np.abs(qp.A_vals))
            self.con_scale[:] = np.minimum(1.0, 100.0 / rowmax)
            H0 = sp.coo_matrix((qp.H_vals, (qp.H_rows, qp.H_cols)), shape=(nx, nx)).tocsr()
            g0 = H0 @ x0 + sp.tril(H0, -1).T @ x0 + qp.c
            gmax = float(np.max(np.abs(g0))) if nx else 0.0
            self.obj_scale = min(1.0, 100.0 / gmax) if gmax > 0 else 1.0
            self.y[:] = self.y / self.con_scale
            self.rhs[:] = self.rhs * self.con_scale
            cs = self.con_scale[ic.ind_ineq]
            self.x[nx:] *= cs; self.xl[nx:] *= cs; self.xu[nx:] *= cs

        # MadNLP.initialize!(kkt)  (normalkkt.jl:136-147) + init_regularization! (kernels.jl:380-408)
        self.reg[:] = 1.0; self.pr_diag[:] = 1.0; self.du_diag[:] = 0.0
        self.l_lower[:] = 0.0; self.u_lower[:] = 0.0; self.l_diag[:] = 1.0; self.u_diag[:] = 1.0
        self.scaling_factor[:] = 1.0
        kind = opt.regularization[0]
        self.del_w = 1.0
        self.del_c = 0.0 if kind == "none" else opt.regularization[2]
        self._reg_state = list(opt.regularization)

        # callbacks: jac / hess once (solver.jl:166-170), compress_* (normalkkt.jl:149-160)
        self._cost = np.concatenate([qp.c * self.obj_scale, np.zeros(ns)])
        self.jac_V[:qp.nnzj] = qp.A_vals * self.con_scale[qp.A_rows]
        self.jac_V[qp.nnzj:] = -1.0
        self.AT_nzval[:] = self.jac_V[self.A_csr_map - 1]
        self._Amat = self._A()
        self._Hs = None
        if self._k2:
            self.hess_V[:] = qp.H_vals * self.obj_scale
            self.aug_V[:] = 0.0
            self.aug_V[self._off_hess:self._off_hess + qp.nnzh] = self.hess_V
            self.aug_V[self._off_jac:self._off_jac + self.jac_V.size] = self.jac_V
        else:
            self.hess_V = np.zeros(0)
        self.obj_val = self.eval_f()
        self.eval_grad_f()
        self.eval_cons()
        self.norm_b = float(np.max(np.abs(self.rhs))) if m else 0.0
        self.norm_c = float(np.max(np.abs(self.f))) if n else 0.0
        self.init_starting_point()
        self.mu = opt.mu_init

    # ---- KKT system operations ---------------------------------------------------------------
    def build_kkt(self):
        if self._scaled:
            # MadNLP.build_kkt!(::ScaledSparseKKTSystem) [RECALLED] = _build_scale_augmented_system_coo!
            # (/root/reference/scripts/cuda_wrapper.jl:90-116, in repo) + transfer! (:17-28)
            n, m = self.n, self.m
            src = self.aug_V
            src[:n] = self.pr_diag
            src[self._off_du:] = self.du_diag
            dest = scale_augmented_system_coo(self._aug_I, self._aug_J, src, self.scaling_factor, n, m)
            transfer(self.aug_nzval, dest, self.aug_csc_map)
        elif self._k2:
            n = self.n
            self.aug_V[:n] = self.pr_diag                       # views into the COO values [RECALLED]
            self.aug_V[self._off_du:] = self.du_diag
            transfer(self.aug_nzval, self.aug_V, self.aug_csc_map)
        else:
            # normalkkt.jl:166-180; du_diag is NOT added (SURVEY note B.2)
            D = self.buffer_n
            D[:] = 1.0 / self.pr_diag
            assemble_normal_system(self.m, self.n, self.Ap, self.Aj, self.AT_nzval, self.aug_colptr,
                                   self.aug_rowval, self.aug_nzval, D)

    def factorize_wrapper(self):
        """MadNLP.factorize_wrapper! = build_kkt! + factorize! [RECALLED]."""
        self.build_kkt()
        t0 = time.perf_counter()
        self.linear_solver.factorize()
        self.t_factor += time.perf_counter() - t0
        self.n_factor += 1

    def kkt_solve(self, w):
        """MadNLP.solve!(kkt, w): K2 [RECALLED, SURVEY 8a K4'] / Normal (normalkkt.jl:182-205)."""
        wx, wy, wzl, wzu = self._views(w)
        if self._scaled:
            return self._kkt_solve_scaled(w)
        # reduce_rhs!
        wx[self.ind_lb] -= wzl / self.l_diag
        wx[self.ind_ub] -= wzu / self.u_diag
        if self._k2:
            self.linear_solver.solve(w[:self.n + self.m])
        else:
            r1, r2, S = self.buffer_n, self.buffer_m, self.pr_diag
            r1[:] = wx / S
            r2[:] = wy
            r2[:] = self._Amat @ r1 - r2
            self.linear_solver.solve(r2)
            wy[:] = r2
            r1[:] = wx
            r1[:] = r1 - self._Amat.T @ wy
            wx[:] = r1 / S
        # finish_aug_solve!
        wzl[:] = (-wzl + self.l_lower * wx[self.ind_lb]) / self.l_diag
        wzu[:] = (wzu - self.u_lower * wx[self.ind_ub]) / self.u_diag
        return w

    def _kkt_solve_scaled(self, w):
        """MadNLP.solve!(::ScaledSparseKKTSystem, w) [MadNLP 0.8.7, un-vendored; RECALLED].  The algebra is forced: the
        K2.5 matrix is T*K2*T with T = diag(scaling_factor, I) (build_kkt above, in repo), the unreduced rows are
        zl*dx + (x-xl)*dzl = wzl and zu*dx - (xu-x)*dzu = wzu (kernels.jl:38-39,58-59 with l_diag = x-xl, u_diag = xu-x
        of kernels.jl:158-159), so  rhs = s*wx + wzl*s/l_diag + wzu*s/u_diag,  dx = s*dx~,  and the back-substitution
        below.  s/l_diag is evaluated as sqrt(u_diag)/sqrt(l_diag) (1/sqrt(l_diag) without an upper bound)."""
        wx, wy, wzl, wzu = self._views(w)
        n = self.n
        sa = np.ones(n); sb = np.ones(n)
        sa[self.ind_lb] = np.sqrt(self.l_diag)
        sb[self.ind_ub] = np.sqrt(self.u_diag)
        r = wx * self.scaling_factor
        r[self.ind_lb] += wzl * sb[self.ind_lb] / sa[self.ind_lb]
        r[self.ind_ub] += wzu * sa[self.ind_ub] / sb[self.ind_ub]
        wx[:] = r
        self.linear_solver.solve(w[:n + self.m])
        wx *= self.scaling_factor
        wzl[:] = (wzl - self.l_lower * wx[self.ind_lb]) / self.l_diag
        wzu[:] = (self.u_lower * wx[self.ind_ub] - wzu) / self.u_diag
        return w

    def kkt_mul(self, w, v, alpha=1.0, beta=0.0):
        """mul!(w, kkt, v, alpha, beta): normalkkt.jl:207-219 + _kktmul! [RECALLED, SURVEY K5]."""
        wx, wy, wzl, wzu = self._views(w)
        vx, vy, vzl, vzu = self._views(v)
        hx = self._Hsym() @ vx if self._k2 and self.qp.nnzh else 0.0
        wx[:] = alpha * (hx + self._Amat.T @ vy) + beta * wx
        wy[:] = alpha * (self._Amat @ vx) + beta * wy
        wx += alpha * self.reg * vx
        wy += alpha * self.du_diag * vy
        wx[self.ind_lb] -= alpha * vzl
        wx[self.ind_ub] += alpha * vzu
        sg = -1.0 if self._scaled else 1.0                  # K2.5 stores l_diag = x - xl, u_diag = xu - x (kernels.jl:158-159)
        wzl[:] = beta * wzl + alpha * (vx[self.ind_lb] * self.l_lower - vzl * (sg * self.l_diag))
        wzu[:] = beta * wzu + alpha * (vx[self.ind_ub] * self.u_lower + vzu * (sg * self.u_diag))
        return w

    def solve_system(self, d, p):
        """/root/reference/src/linear_solver.jl:19-45."""
        d[:] = p
        self.kkt_solve(d)
        w = self._w1
        w[:] = p
        self.kkt_mul(w, d, -1.0, 1.0)
        norm_w = float(np.max(np.abs(w))) if w.size else 0.0
        norm_p = float(np.max(np.abs(p))) if p.size else 0.0
        ratio = norm_w / max(1.0, norm_p)
        self.last_residual_ratio = ratio
        if np.isnan(ratio) or (self.opt.check_residual and ratio > self.opt.tol_linear_solve):
            raise SolveException()
        return d

    # ---- kernels.jl --------------------------------------------------------------------------
    def set_aug_diagonal_reg(self):
        """/root/reference/src/kernels.jl:128-146; ScaledSparseKKTSystem method: kernels.jl:149-165."""
        self.reg[:] = self.del_w
        self.du_diag[:] = self.del_c
        if self._scaled:
            self.l_diag[:] = self.x[self.ind_lb] - self.xl[self.ind_lb]     # (X - Xl), kernels.jl:158
            self.u_diag[:] = self.xu[self.ind_ub] - self.x[self.ind_ub]     # (Xu - X), kernels.jl:159
            self.l_lower[:] = self.zl[self.ind_lb]
            self.u_lower[:] = self.zu[self.ind_ub]
            # MadNLP._set_aug_diagonal!(::ScaledSparseKKTSystem) [MadNLP 0.8.7, un-vendored; RECALLED]
            xlzu = np.zeros(self.n); xuzl = np.zeros(self.n)
            xlzu[self.ind_ub] = self.u_lower
            xlzu[self.ind_lb] *= self.l_diag
            xuzl[self.ind_lb] = self.l_lower
            xuzl[self.ind_ub] *= self.u_diag
            self.pr_diag[:] = xlzu + xuzl
            self.scaling_factor[:] = 1.0
            self.scaling_factor[self.ind_lb] *= np.sqrt(self.l_diag)
            self.scaling_factor[self.ind_ub] *= np.sqrt(self.u_diag)
            self.pr_diag += self.reg * (self.scaling_factor * self.scaling_factor)
            return
        self.l_diag[:] = self.xl[self.ind_lb] - self.x[self.ind_lb]
        self.u_diag[:] = self.x[self.ind_ub] - self.xu[self.ind_ub]
        self.l_lower[:] = self.zl[self.ind_lb]
        self.u_lower[:] = self.zu[self.ind_ub]
        self.pr_diag[:] = self.reg
        self.pr_diag[self.ind_lb] -= self.l_lower / self.l_diag
        self.pr_diag[self.ind_ub] -= self.u_lower / self.u_diag

    def factorize_regularized_system(self):
        """/root/reference/src/linear_solver.jl:6-17."""
        for _ in range(3):
            self.set_aug_diagonal_reg()
            self.factorize_wrapper()
            if self.linear_solver.factorized:
                break
            self.del_w *= 100.0
            self.del_c *= 100.0

    def set_initial_primal_rhs(self):
        """kernels.jl:1-9."""
        self.p[:] = 0.0
        self._views(self.p)[1][:] = -self.c

    def set_initial_dual_rhs(self):
        """kernels.jl:11-19."""
        self.p[:] = 0.0
        self._views(self.p)[0][:] = -self.f

    def set_predictive_rhs(self):
        """kernels.jl:21-41."""
        px, py, pzl, pzu = self._views(self.p)
        self.p[:] = 0.0
        px[:] = -self.f + self.zl - self.zu - self.jacl
        py[:] = -self.c
        pzl[:] = (self.xl[self.ind_lb] - self.x[self.ind_lb]) * self.zl[self.ind_lb]
        pzu[:] = (self.xu[self.ind_ub] - self.x[self.ind_ub]) * self.zu[self.ind_ub]

    def set_correction_rhs(self, mu):
        """kernels.jl:43-61."""
        px, py, pzl, pzu = self._views(self.p)
        px[:] = -self.f + self.zl - self.zu - self.jacl
        py[:] = -self.c
        pzl[:] = (self.xl[self.ind_lb] - self.x[self.ind_lb]) * self.zl[self.ind_lb] + mu - self.correction_lb
        pzu[:] = (self.xu[self.ind_ub] - self.x[self.ind_ub]) * self.zu[self.ind_ub] - mu - self.correction_ub

    def get_correction(self):
        """kernels.jl:63-75."""
        dx, _, dzl, dzu = self._views(self.d)
        self.correction_lb[:] = dx[self.ind_lb] * dzl
        self.correction_ub[:] = dx[self.ind_ub] * dzu

    def set_extra_correction(self, alpha_p, alpha_d, bmin, bmax, mu):
        """kernels.jl:78-126."""
        dx, _, dzl, dzu = self._views(self.d)
        tmin, tmax = bmin * mu, bmax * mu
        x_ = self.x[self.ind_lb] + alpha_p * dx[self.ind_lb] - self.xl[self.ind_lb]
        z_ = self.zl[self.ind_lb] + alpha_d * dzl
        v = x_ * z_
        delta = np.where(v < tmin, tmin - v, np.where(v > tmax, tmax - v, 0.0))
        self.correction_lb[:] = self.correction_lb - delta
        x_ = self.xu[self.ind_ub] - alpha_p * dx[self.ind_ub] - self.x[self.ind_ub]
        z_ = self.zu[self.ind_ub] + alpha_d * dzu
        v = x_ * z_
        delta = np.where(v < tmin, tmin - v, np.where(v > tmax, tmax - v, 0.0))
        self.correction_ub[:] = self.correction_ub + delta

    def get_complementarity_measure(self):
        """kernels.jl:171-190."""
        if self.nlb + self.nub == 0:
            return 0.0
        l = float(np.sum((self.x[self.ind_lb] - self.xl[self.ind_lb]) * self.zl[self.ind_lb]))
        u = float(np.sum((self.xu[self.ind_ub] - self.x[self.ind_ub]) * self.zu[self.ind_ub]))
        return (l + u) / (self.nlb + self.nub)

    def get_affine_complementarity_measure(self, alpha_p, alpha_d):
        """kernels.jl:192-224."""
        if self.nlb + self.nub == 0:
            return 0.0
        dx, _, dzl, dzu = self._views(self.d)
        lb, ub = self.ind_lb, self.ind_ub
        l = float(np.sum(((self.x[lb] + alpha_p * dx[lb]) - self.xl[lb]) * (self.zl[lb] + alpha_d * dzl)))
        u = float(np.sum((self.xu[ub] - (self.x[ub] + alpha_p * dx[ub])) * (self.zu[ub] + alpha_d * dzu)))
        return (l + u) / (self.nlb + self.nub)

    def update_barrier(self, mu_affine):
        """kernels.jl:226-236 (has_inequalities per SURVEY note B.1: any finite bound)."""
        # ind_llb / ind_uub fields receive ind_lb / ind_ub (structure.jl:170-171 vs 48-51)
        has_inequalities = (self.nlb + self.nub) > 0
        mu_curr = self.get_complementarity_measure()
        if has_inequalities:
            r = mu_affine / mu_curr
            sigma = min(max(r * r * r, 1e-6), 10.0)          # Julia lowers ^3 with a literal exponent to x*x*x (Base.literal_pow)
        else:
            sigma = 1.0
        self.mu = max(self.opt.mu_min, sigma * mu_curr)
        return mu_curr

    @staticmethod
    def _argmin_fold(vals):
        """mapreduce with init (1.0, 0) and op (a, b) -> a[1] < b[1] ? a : b (kernels.jl:243-251):
        value min(1, min vals); on ties the later index wins; 1-based index, 0 if none."""
        best, idx = 1.0, 0
        if vals.size:
            m = float(np.min(vals))
            if m <= best:          # equal also replaces (strict '<' keeps the left only if smaller)
                best = m
                idx = int(np.nonzero(vals == m)[0][-1]) + 1
        return best, idx

    def get_alpha_max_primal(self, tau):
        """kernels.jl:242-264."""
        dx = self._views(self.d)[0]
        lb, ub = self.ind_lb, self.ind_ub
        with np.errstate(divide="ignore", invalid="ignore"):
            dxl, dxu = dx[lb], dx[ub]
            vl = np.where(dxl < 0, (-self.x[lb] + self.xl[lb]) * tau / dxl, np.inf)
            vu = np.where(dxu > 0, (-self.x[ub] + self.xu[ub]) * tau / dxu, np.inf)
        (al, il), (au, iu) = self._argmin_fold(vl), self._argmin_fold(vu)
        return al, au, il, iu

    def get_alpha_max_dual(self, tau):
        """kernels.jl:266-288."""
        _, _, dzl, dzu = self._views(self.d)
        zl_r, zu_r = self.zl[self.ind_lb], self.zu[self.ind_ub]
        with np.errstate(divide="ignore", invalid="ignore"):
            vl = np.where(dzl < 0, (-zl_r) * tau / dzl, np.inf)
            vu = np.where((dzu < 0) & (zu_r + dzu < 0), (-zu_r) * tau / dzu, np.inf)
        (al, il), (au, iu) = self._argmin_fold(vl), self._argmin_fold(vu)
        return al, au, il, iu

    def get_fraction_to_boundary_step(self, tau):
        """kernels.jl:290-305."""
        axl, axu, _, _ = self.get_alpha_max_primal(tau)
        azl, azu, _, _ = self.get_alpha_max_dual(tau)
        return min(axl, axu), min(azl, azu)

    def update_step(self):
        """kernels.jl:307-374."""
        rule = self.opt.step_rule
        if rule[0] == "conservative":
            self.alpha_p, self.alpha_d = self.get_fraction_to_boundary_step(rule[1])
        elif rule[0] == "adaptive":
            tau = max(1 - self.mu, rule[1])
            self.alpha_p, self.alpha_d = self.get_fraction_to_boundary_step(tau)
        elif rule[0] == "mehrotra":
            gamma_f = rule[1]
            gamma_a = 1.0 / (1.0 - gamma_f)
            dx, _, d_zl, d_zu = self._views(self.d)
            lb, ub = self.ind_lb, self.ind_ub
            axl, axu, i_xl, i_xu = self.get_alpha_max_primal(1.0)
            azl, azu, i_zl, i_zu = self.get_alpha_max_dual(1.0)
            max_ap, max_ad = min(axl, axu), min(azl, azu)
            mu_full = self.get_affine_complementarity_measure(max_ap, max_ad) / gamma_a
            alpha_p = alpha_d = 1.0
            if max_ap < 1.0:
                if axl <= axu:
                    i = i_xl - 1
                    tmp = mu_full / (self.zl[lb][i] + max_ad * d_zl[i])
                    alpha_p = (self.x[lb][i] - self.xl[lb][i] - tmp) / (-dx[lb][i])
                else:
                    i = i_xu - 1
                    tmp = mu_full / (self.zu[ub][i] + max_ad * d_zu[i])
                    alpha_p = (self.xu[ub][i] - self.x[ub][i] - tmp) / (dx[ub][i])
            if max_ad < 1.0:
                if azl <= azu:
                    i = i_zl - 1
                    tmp = mu_full / (self.x[lb][i] + max_ap * dx[lb][i] - self.xl[lb][i])
                    alpha_d = -(self.zl[lb][i] - tmp) / d_zl[i]
                else:
                    i = i_zu - 1
                    tmp = mu_full / (self.xu[ub][i] - self.x[ub][i] - max_ap * dx[ub][i])
                    alpha_d = -(self.zu[ub][i] - tmp) / d_zu[i]
            self.alpha_p = max(alpha_p, gamma_f * max_ap)
            self.alpha_d = max(alpha_d, gamma_f * max_ad)
        else:
            raise ValueError(rule)

    def update_regularization(self):
        """kernels.jl:386-417."""
        r = self._reg_state
        if r[0] == "none":
            self.del_w = 0.0; self.del_c = 0.0
        elif r[0] == "fixed":
            self.del_w = r[1]; self.del_c = r[2]
        elif r[0] == "adaptive":
            r[1] = max(r[1] / 10.0, r[3])
            r[2] = min(r[2] / 10.0, -r[3])
            self.del_w = r[1]; self.del_c = r[2]

    def get_optimality_gap(self):
        """kernels.jl:435-446 -> MadNLP.get_inf_compl(..., mu=0, sc=1) [RECALLED]."""
        a = np.max(np.abs((self.x[self.ind_lb] - self.xl[self.ind_lb]) * self.zl[self.ind_lb])) if self.nlb else 0.0
        b = np.max(np.abs((self.xu[self.ind_ub] - self.x[self.ind_ub]) * self.zu[self.ind_ub])) if self.nub else 0.0
        return float(max(a, b))

    # ---- solver.jl ---------------------------------------------------------------------------
    def init_starting_point(self):
        """/root/reference/src/solver.jl:6-125."""
        lb_i, ub_i = self.ind_lb, self.ind_ub
        x = self.x
        l, u = self.xl, self.xu
        self.reg[:] = self.del_w
        self.pr_diag[:] = self.del_w
        self.du_diag[:] = self.del_c
        self.factorize_wrapper()
        self.set_initial_primal_rhs()
        self.solve_system(self.d, self.p)
        x += self._views(self.d)[0]
        self.set_initial_dual_rhs()
        self.solve_system(self.d, self.p)
        self.y[:] = self._views(self.d)[1]
        res = self.jacl
        res[:] = self._Amat.T @ self.y
        res += self.f
        fl, fu = np.isfinite(l), np.isfinite(u)
        self.zl[:] = np.where(fl & fu, 0.5 * res, np.where(fl, res, self.zl))
        self.zu[:] = np.where(fl & fu, -0.5 * res, np.where(fu, -res, self.zu))
        xl, xu = x[lb_i], x[ub_i]
        lb, ub = l[lb_i], u[ub_i]
        zl, zu = self.zl[lb_i], self.zu[ub_i]
        delta_x = max(0.0, -1.5 * min(0.0, np.min(xl - lb, initial=0.0)), -1.5 * min(0.0, np.min(ub - xu, initial=0.0)))
        delta_s = max(0.0, -1.5 * min(0.0, np.min(zl, initial=0.0)), -1.5 * min(0.0, np.min(zu, initial=0.0)))
        # NOTE: xl and xu are index views of the SAME x; the two updates are applied in turn (l.80-81)
        x[lb_i] = x[lb_i] + delta_x
        x[ub_i] = x[ub_i] - delta_x
        self.zl[lb_i] += 1.0 + delta_s
        self.zu[ub_i] += 1.0 + delta_s
        xl, xu, zl, zu = x[lb_i], x[ub_i], self.zl[lb_i], self.zu[ub_i]
        mu = 0.0
        if zl.size:
            mu += float(xl @ zl) - float(lb @ zl)
        if zu.size:
            mu += float(ub @ zu) - float(xu @ zu)
        delta_x2 = mu / (2 * (float(zl.sum()) + float(zu.sum())))
        delta_s2 = mu / (2 * (float((xl - lb).sum()) + float((ub - xu).sum())))
        x[lb_i] += delta_x2
        x[ub_i] -= delta_x2
        self.zl[lb_i] += delta_s2
        self.zu[ub_i] += delta_s2
        kappa = self.opt.bound_fac
        with np.errstate(invalid="ignore"):
            pl = np.minimum(kappa * np.maximum(1.0, l), kappa * (u - l))
            pu = np.minimum(kappa * np.maximum(1.0, u), kappa * (u - l))
            x[:] = np.where(x < l, l + pl, np.where(u < x, u - pu, x))
        assert np.all(self.zl[lb_i] > 0.0) and np.all(self.zu[ub_i] > 0.0)
        assert np.all(x[lb_i] > l[lb_i]) and np.all(x[ub_i] < u[ub_i])

    def gondzio_correction_direction(self, mu_curr, max_ncorr):
        """/root/reference/src/solver.jl:200-251."""
        delta, bmin, bmax, tau = 0.1, 0.1, 10.0, 0.995
        dp = self._w2
        alpha_p, alpha_d = self.get_fraction_to_boundary_step(tau)
        for _ in range(max_ncorr):
            ta_p = min(alpha_p + delta, 1.0)
            ta_d = min(alpha_d + delta, 1.0)
            ga = self.get_affine_complementarity_measure(ta_p, ta_d)
            g = mu_curr
            mu = (ga / g) ** 2 * ga
            self.set_extra_correction(ta_p, ta_d, bmin, bmax, mu)
            self.set_correction_rhs(mu)
            dp[:] = self.d
            self.solve_system(self.d, self.p)
            hp, hd = self.get_fraction_to_boundary_step(tau)
            if hp < 1.005 * alpha_p or hd < 1.005 * alpha_d:
                self.d[:] = dp
                break
            alpha_p, alpha_d = hp, hd
        return alpha_p, alpha_d

    def adjust_boundary(self):
        """MadNLP.adjust_boundary! [RECALLED], call site solver.jl:342."""
        eps = np.finfo(float).eps
        c1, c2 = eps * self.mu, eps ** 0.75
        lb, ub = self.ind_lb, self.ind_ub
        x_lr, xl_r = self.x[lb], self.xl[lb]
        self.xl[lb] = np.where(x_lr - xl_r < c1, xl_r - c2 * np.maximum(1.0, np.abs(x_lr)), xl_r)
        x_ur, xu_r = self.x[ub], self.xu[ub]
        self.xu[ub] = np.where(xu_r - x_ur < c1, xu_r + c2 * np.maximum(1.0, np.abs(x_ur)), xu_r)

    def mpc(self):
        """/root/reference/src/solver.jl:254-345."""
        opt = self.opt
        while True:
            self.jacl[:] = self._Amat.T @ self.y
            self.inf_pr = (float(np.max(np.abs(self.c))) if self.m else 0.0) / max(1.0, self.norm_b)
            self.inf_du = float(np.max(np.abs(self.f - self.zl + self.zu + self.jacl))) / max(1.0, self.norm_c)
            self.inf_compl = self.get_optimality_gap() / max(1.0, self.norm_c)
            self.trace.append(dict(k=self.k, obj=self.obj_val / self.obj_scale, inf_pr=self.inf_pr,
                                   inf_du=self.inf_du, inf_compl=self.inf_compl, mu=self.mu,
                                   alpha_p=self.alpha_p, alpha_d=self.alpha_d))
            if max(self.inf_pr, self.inf_du, self.inf_compl) <= opt.tol:
                return SOLVE_SUCCEEDED
            if self.k >= opt.max_iter:
                return MAXIMUM_ITERATIONS_EXCEEDED
            self.update_regularization()
            self.factorize_regularized_system()
            # prediction
            self.set_predictive_rhs()
            self.solve_system(self.d, self.p)
            a_aff_p, a_aff_d = self.get_fraction_to_boundary_step(1.0)
            mu_affine = self.get_affine_complementarity_measure(a_aff_p, a_aff_d)
            self.get_correction()
            mu_curr = self.update_barrier(mu_affine)
            # correction
            self.set_correction_rhs(self.mu)
            self.solve_system(self.d, self.p)
            if opt.max_ncorr > 0:
                self.gondzio_correction_direction(mu_curr, opt.max_ncorr)
            self.update_step()
            dx, dy, dzl, dzu = self._views(self.d)
            self.x += self.alpha_p * dx
            self.y += self.alpha_d * dy
            self.zl[self.ind_lb] += self.alpha_d * dzl
            self.zu[self.ind_ub] += self.alpha_d * dzu
            self.obj_val = self.eval_f()
            self.eval_cons()
            self.eval_grad_f()
            self.adjust_boundary()
            self.k += 1

    def solve(self):
        """/root/reference/src/solver.jl:347-403."""
        t0 = time.perf_counter()
        try:
            self.initialize()
            self.t_init = time.perf_counter() - t0
            self.status = self.mpc()
        except SolveException:
            self.status = ERROR_IN_STEP_COMPUTATION
        except (AssertionError, FloatingPointError, ZeroDivisionError):
            self.status = INTERNAL_ERROR
        self.total_time = time.perf_counter() - t0
        return self

    # ---- results in the user's (unscaled) space ------------------------------------------------
    @property
    def objective(self):
        return self.obj_val / self.obj_scale

    @property
    def solution(self):
        if self._fixed.size == 0:
            return self.x[:self.nx].copy()
        full = np.empty(self.qp_full.nvar)
        full[self._free] = self.x[:self.nx]
        full[self._fixed] = self._x_fixed
        return full

    @property
    def multipliers(self):
        return self.y * self.con_scale / self.obj_scale
