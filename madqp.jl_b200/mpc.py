"""Host-side mirror of MadIPM's solver interface on top of the B200 backend.

In the target deployment the Mehrotra predictor-corrector control
(/root/reference/src/solver.jl) stays in Julia and reaches the GPU through the
``ccall`` shim of INTEGRATION.md.  Julia is not available in this environment,
so this module is the stand-in caller: ``MPCSolver`` has the reference's option
names, status names and call structure (``initialize!``, ``mpc!``, ``solve!``),
and every numeric step is one call through the C ABI (``_lib.Handle``).  No
numerical work is done here beyond scalars; nothing falls back to the CPU.

One ``MPCSolver`` drives ``batch`` instances that share a sparsity pattern
(batch = 1 for a single solve; the C5 configuration uses batch > 1).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import Handle, MadIPMError
from .preprocess import make_parameter

# MadNLP.Status names used by the reference (solver.jl:279-283, 359-390)
SOLVE_SUCCEEDED = "SOLVE_SUCCEEDED"
MAXIMUM_ITERATIONS_EXCEEDED = "MAXIMUM_ITERATIONS_EXCEEDED"
ERROR_IN_STEP_COMPUTATION = "ERROR_IN_STEP_COMPUTATION"
INTERNAL_ERROR = "INTERNAL_ERROR"
INITIAL = "INITIAL"


class SolveException(Exception):
    """MadNLP.SolveException, /root/reference/src/linear_solver.jl:41-43."""


# ---- option types of /root/reference/src/utils.jl:17-48 ------------------------------------------
@dataclass
class ConservativeStep:
    tau: float = 0.995


@dataclass
class AdaptiveStep:
    tau_min: float = 0.99


@dataclass
class MehrotraAdaptiveStep:
    gamma_f: float = 0.99


@dataclass
class NoRegularization:
    pass


@dataclass
class FixedRegularization:
    delta_p: float = 1e-8
    delta_d: float = 0.0


@dataclass
class AdaptiveRegularization:
    delta_p: float = 1e-8
    delta_d: float = -1e-9
    delta_min: float = 1e-9


_KKT_KINDS = {"K2": _lib.KKT_K2, "NORMAL": _lib.KKT_NORMAL, "K2.5": _lib.KKT_K2_SCALED}


@dataclass
class IPMOptions:
    """/root/reference/src/utils.jl:69-103 (fields that reach the hot path)."""
    tol: float = 1e-8
    kkt_system: str = "K2"               # "K2" = MadNLP.SparseKKTSystem (default), "NORMAL" = MadIPM.NormalKKTSystem,
                                         # "K2.5" = MadNLP.ScaledSparseKKTSystem (kernels.jl:149-165, runtests.jl:102-115)
    max_iter: int = 3000
    scaling: bool = True
    bound_push: float = 1e-2
    bound_fac: float = 1e-2
    bound_relax_factor: float = 1e-8
    regularization: object = field(default_factory=lambda: FixedRegularization(1e-8, 0.0))
    step_rule: object = field(default_factory=lambda: AdaptiveStep(0.99))
    max_ncorr: int = 0
    mu_init: float = 1e-1
    mu_min: float = 1e-11
    tol_linear_solve: float = 1e-8
    check_residual: bool = False
    # backend options (the linear-solver options of the reference's `opt_linear_solver`)
    ordering: int = _lib.ORDER_ND
    max_width: int = 64
    block_panel: int = 0            # 0 = library default (csrc/symbolic.hpp)
    fused_phases: bool = True       # 3 host synchronisations per iteration (madipm_phase_*) instead of 7; same arithmetic
    # Deviation from linear_solver.jl:14-15, acts only after a breakdown (where the reference's unpivoted LDLSolver path fails
    # as well): the retry multiplies (del_w, del_c) by 100, which leaves del_c = 0 (the reference's default delta_d) at 0 for
    # ever; with the kick the first retry only sets del_c = -del_w, so dependent constraint rows get a usable pivot.
    dual_regularization_kick: bool = True
    # The predictor solve's residual ratio (linear_solver.jl:28-36, computed anyway) above this value counts as a breakdown of
    # the factorisation as well: exactly dependent rows leave pivots AND their columns at roundoff size, which neither pivot
    # test can see, but the direction is garbage and the residual says so (healthy solves: 1e-10 ... 1e-6).  0 = off.
    residual_breakdown_tol: float = 1e-3
    growth_limit: float = 1e13      # element-growth breakdown test of the unpivoted LDL' (madipm_set_pivot_checks); 0 = off
    tiny_pivot: float = 1e-13       # |pivot| at or below it is a breakdown as well; 0 = only exact zeros (the reference)
    fused_iteration: bool = True    # ... and a whole iteration behind ONE synchronisation (madipm_phase_iteration) when max_ncorr = 0
    fused_factor_forward: bool = True   # ... and the predictor's forward sweep inside the factor launch (madipm_phase_factor_predictor)
    device_prepare: bool = True     # initialize! / set_scaling! on the device (madipm_load_problem) instead of numpy
    relax_small: int = 0            # relaxed amalgamation (0 / None = library defaults)
    relax_zeros: tuple = None
    nd_leaf: int = 0
    device: int = 0


@dataclass
class IndexConstraints:
    ind_eq: np.ndarray
    ind_ineq: np.ndarray
    ind_lb: np.ndarray
    ind_ub: np.ndarray


def get_index_constraints(lvar, uvar, lcon, ucon):
    """MadNLP.get_index_constraints (call site /root/reference/src/structure.jl:95-102); 0-based."""
    if np.any(lvar == uvar):
        raise NotImplementedError("fixed variables reach get_index_constraints: apply preprocess.make_parameter first "
                                  "(MPCSolver does, fixed_variable_treatment = MakeParameter)")
    ind_eq = np.nonzero(lcon == ucon)[0]
    ind_ineq = np.nonzero(lcon != ucon)[0]
    xl = np.concatenate([lvar, lcon[ind_ineq]])
    xu = np.concatenate([uvar, ucon[ind_ineq]])
    return IndexConstraints(ind_eq, ind_ineq, np.nonzero(xl > -np.inf)[0], np.nonzero(xu < np.inf)[0])


def _push_inside(x, l, u, bound_push, bound_fac):
    """MadNLP.initialize_variables! (Ipopt's projection of x0 into the interior of [l, u])."""
    fl, fu = np.isfinite(l), np.isfinite(u)
    with np.errstate(invalid="ignore"):
        pl = bound_push * np.maximum(1.0, np.abs(l))
        pu = bound_push * np.maximum(1.0, np.abs(u))
        width = bound_fac * (u - l)
        inside = np.minimum(u - np.minimum(pu, width), np.maximum(l + np.minimum(pl, width), x))
        return np.where(fl & fu, inside, np.where(fl, np.maximum(l + pl, x), np.where(fu, np.minimum(u - pu, x), x)))


@dataclass
class ExecutionStats:
    status: list
    iter: np.ndarray
    objective: np.ndarray
    solution: np.ndarray
    multipliers: np.ndarray
    inf_pr: np.ndarray
    inf_du: np.ndarray
    inf_compl: np.ndarray
    total_time: float
    init_time: float


class B200KKTSystem:
    """The AbstractKKTSystem + linear-solver pair as one object
    (method list of /root/reference/src/KKT/normalkkt.jl:128-220)."""

    def __init__(self, handle: Handle, kind: str):
        self.h = handle
        self.kind = kind

    def initialize(self):                      # MadNLP.initialize!(kkt)
        self.h.kkt_initialize()

    def build_kkt(self):                       # MadNLP.build_kkt!
        self.h.build_kkt()

    def factorize(self):                       # MadNLP.factorize!(kkt.linear_solver)
        return self.h.factorize()

    def is_factorized(self):                   # MadIPM.is_factorized
        return self.h.is_factorized()

    def solve(self, w="D"):                    # MadNLP.solve!(kkt, w)
        self.h.kkt_solve(w)

    def mul(self, w, v, alpha=1.0, beta=0.0):  # LinearAlgebra.mul!(w, kkt, v, alpha, beta)
        self.h.kkt_mul(w, v, alpha, beta)

    def jtprod(self, out, inp):                # MadNLP.jtprod!
        self.h.spmv(_lib.OP_AT, out, inp)

    def is_inertia_correct(self, n_pos, n_zero, n_neg, n, m):
        if self.kind == "NORMAL":
            return n_zero == 0 and n_pos == m          # normalkkt.jl:132-134
        return n_zero == 0 and n_pos == n and n_neg == m


class MPCSolver:
    """MadIPM.MPCSolver(qp; kwargs...) (/root/reference/src/structure.jl:77-176)."""

    def __init__(self, qp, **kwargs):
        self.qps = list(qp) if isinstance(qp, (list, tuple)) else [qp]
        # fixed_variable_treatment = MakeParameter (utils.jl:81): variables with lvar == uvar leave the model before the hot path
        self.nx_full = self.qps[0].nvar
        reduced = [make_parameter(q) for q in self.qps]
        self._fixed, self._free = reduced[0][1], reduced[0][2]
        if any(not np.array_equal(r[1], self._fixed) for r in reduced[1:]):
            raise ValueError("batched instances must fix the same variables")
        self._x_fixed = np.stack([r[3] for r in reduced])
        self.qps = [r[0] for r in reduced]
        q0 = self.qps[0]
        self.B = len(self.qps)
        self.opt = IPMOptions(**kwargs)
        opt = self.opt
        t0 = time.perf_counter()
        ic = get_index_constraints(q0.lvar, q0.uvar, q0.lcon, q0.ucon)
        for q in self.qps[1:]:
            jc = get_index_constraints(q.lvar, q.uvar, q.lcon, q.ucon)
            if not (np.array_equal(ic.ind_ineq, jc.ind_ineq) and np.array_equal(ic.ind_lb, jc.ind_lb)
                    and np.array_equal(ic.ind_ub, jc.ind_ub) and np.array_equal(q.A_rows, q0.A_rows)
                    and np.array_equal(q.A_cols, q0.A_cols) and np.array_equal(q.H_rows, q0.H_rows)):
                raise ValueError("batched instances must share one sparsity pattern and bound structure")
        self.ic = ic
        self.nx, self.m = q0.nvar, q0.ncon
        self.ns = ic.ind_ineq.size
        self.n = self.nx + self.ns
        self.nlb, self.nub = ic.ind_lb.size, ic.ind_ub.size
        if opt.kkt_system == "NORMAL" and q0.nnzh > 0:
            raise ValueError("The KKT system NormalKKTSystem supports only linear programs.")   # normalkkt.jl:43-46
        if opt.kkt_system not in _KKT_KINDS:
            raise ValueError(f"unknown kkt_system {opt.kkt_system!r}; expected one of {sorted(_KKT_KINDS)}")
        kind = _KKT_KINDS[opt.kkt_system]
        self.h = Handle(kkt_kind=kind, batch=self.B, device=opt.device, ordering=opt.ordering,
                        max_width=opt.max_width, block_panel=opt.block_panel, nd_leaf=opt.nd_leaf,
                        **({'relax_small': opt.relax_small} if opt.relax_small else {}),
                        **({'relax_zeros': opt.relax_zeros} if opt.relax_zeros else {}))
        if kind != _lib.KKT_NORMAL:
            self.h.set_pattern_k2(self.nx, self.m, q0.H_rows + 1, q0.H_cols + 1, q0.A_rows + 1, q0.A_cols + 1,
                                  ic.ind_ineq + 1)
        else:
            self.h.set_pattern_normal(self.nx, self.m, q0.A_rows + 1, q0.A_cols + 1, ic.ind_ineq + 1)
        self.h.set_bounds_index(ic.ind_lb + 1, ic.ind_ub + 1)
        self.info = self.h.analyze()
        self.kkt = B200KKTSystem(self.h, opt.kkt_system)
        self.analysis_time = time.perf_counter() - t0
        B = self.B
        self.obj_val = np.zeros(B)
        self.obj_scale = np.ones(B)
        self.con_scale = np.ones((B, self.m))
        self.mu = np.zeros(B)
        self.alpha_p = np.zeros(B)
        self.alpha_d = np.zeros(B)
        self.inf_pr = np.zeros(B)
        self.inf_du = np.zeros(B)
        self.inf_compl = np.zeros(B)
        self.norm_b = np.zeros(B)
        self.norm_c = np.zeros(B)
        self.del_w = 0.0
        self.del_c = 0.0
        self.k = 0
        self.iters = np.zeros(B, dtype=np.int64)
        self.status = [INITIAL] * B
        self.active = np.ones(B, dtype=np.int32)
        self.trace = []
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.keep_snapshot = False     # bench.py sets it to time device-resident re-solves

    # ---- solver.jl:127-182 -----------------------------------------------------------------------
    def initialize(self):
        """initialize!(solver): host plumbing + upload of the inputs, then the device-side part."""
        self.prepare_host()
        self.upload()
        self.start()

    def prepare_host(self):
        """MadNLP.initialize!(cb, ...) and set_scaling! (solver.jl:131-159): host-side, per instance.
        Produces the host buffers that `upload` sends to the device."""
        opt, ic, h = self.opt, self.ic, self.h
        B, nx, ns, n, m = self.B, self.nx, self.ns, self.n, self.m
        X = np.zeros((B, n)); XL = np.zeros((B, n)); XU = np.zeros((B, n))
        Y = np.zeros((B, m)); RHS = np.zeros((B, m)); COST = np.zeros((B, n))
        HV = np.zeros((B, self.qps[0].nnzh)); JV = np.zeros((B, self.qps[0].nnzj))
        tol = opt.bound_relax_factor
        for b, qp in enumerate(self.qps):
            # MadNLP.initialize!(cb, x, xl, xu, y, rhs, ind_ineq; tol, bound_push, bound_fac)  (solver.jl:131-142)
            lvar = qp.lvar - np.maximum(1.0, np.abs(qp.lvar)) * tol
            uvar = qp.uvar + np.maximum(1.0, np.abs(qp.uvar)) * tol
            x0 = _push_inside(np.asarray(qp.x0, dtype=float), lvar, uvar, opt.bound_push, opt.bound_fac)
            cons0 = np.bincount(qp.A_rows, weights=qp.A_vals * x0[qp.A_cols], minlength=m)[:m] if m else np.zeros(0)
            ls = qp.lcon[ic.ind_ineq] - np.maximum(1.0, np.abs(qp.lcon[ic.ind_ineq])) * tol
            us = qp.ucon[ic.ind_ineq] + np.maximum(1.0, np.abs(qp.ucon[ic.ind_ineq])) * tol
            s0 = _push_inside(cons0[ic.ind_ineq], ls, us, opt.bound_push, opt.bound_fac)
            X[b] = np.concatenate([x0, s0]); XL[b] = np.concatenate([lvar, ls]); XU[b] = np.concatenate([uvar, us])
            Y[b] = qp.y0
            RHS[b] = np.where(qp.lcon == qp.ucon, qp.lcon, 0.0)
            # MadNLP.set_scaling!(..., 100.0)  (solver.jl:148-159)
            if opt.scaling:
                rowmax = np.ones(m)
                if len(qp.A_rows):                       # max over each row (pattern-only sort, cached per instance)
                    order, starts, present = self._row_groups(b, qp.A_rows)
                    rowmax[present] = np.maximum(1.0, np.maximum.reduceat(np.abs(qp.A_vals)[order], starts))
                self.con_scale[b] = np.minimum(1.0, 100.0 / rowmax)
                g0 = qp.c.astype(float).copy()
                if len(qp.H_rows):
                    off = qp.H_rows != qp.H_cols
                    g0 += np.bincount(qp.H_rows, weights=qp.H_vals * x0[qp.H_cols], minlength=nx)[:nx]
                    g0 += np.bincount(qp.H_cols[off], weights=qp.H_vals[off] * x0[qp.H_rows[off]], minlength=nx)[:nx]
                gmax = float(np.max(np.abs(g0))) if nx else 0.0
                self.obj_scale[b] = min(1.0, 100.0 / gmax) if gmax > 0 else 1.0
                cs = self.con_scale[b]
                Y[b] = Y[b] / cs
                RHS[b] = RHS[b] * cs
                X[b, nx:] *= cs[ic.ind_ineq]; XL[b, nx:] *= cs[ic.ind_ineq]; XU[b, nx:] *= cs[ic.ind_ineq]
            COST[b, :nx] = qp.c * self.obj_scale[b]
            HV[b] = qp.H_vals * self.obj_scale[b]          # eval_lag_hess_wrapper! (solver.jl:170)
            JV[b] = qp.A_vals * self.con_scale[b][qp.A_rows]   # eval_jac_wrapper! (solver.jl:167)
        self._c0 = np.array([qp.c0 for qp in self.qps]) * self.obj_scale
        self._host = dict(X=X, XL=XL, XU=XU, Y=Y, RHS=RHS, COST=COST, HV=HV, JV=JV)
        self._con_scale_on_device = False
        self.norm_b = np.max(np.abs(RHS), axis=1) if m else np.zeros(B)     # solver.jl:173

    def _row_groups(self, b, rows):
        """Stable grouping of a COO row-index array (depends on the pattern only; kept across solves)."""
        cache = self.__dict__.setdefault("_row_group_cache", {})
        if b not in cache:
            order = np.argsort(rows, kind="stable")
            rs = rows[order]
            starts = np.flatnonzero(np.diff(rs, prepend=-1))
            cache[b] = (order, starts, rs[starts])
        return cache[b]

    def load_device(self):
        """prepare_host + upload in one device-side call (madipm_load_problem): the raw model arrays go to the GPU and
        MadNLP.initialize! / set_scaling! / the value compression run there.  prepare_host stays as the host reference."""
        opt, h = self.opt, self.h
        if getattr(self, "_raw", None) is None:
            # the model's arrays, batch-major, in page-locked host memory: marshalled once (they are the caller's inputs and
            # do not change between solves), sent to the device every solve
            self._raw = []
            for f in (lambda q: q.x0, lambda q: q.lvar, lambda q: q.uvar, lambda q: q.y0, lambda q: q.lcon, lambda q: q.ucon,
                      lambda q: q.c, lambda q: q.H_vals, lambda q: q.A_vals):
                first = np.asarray(f(self.qps[0]), dtype=np.float64)
                buf = _lib.PinnedArray((self.B, first.size))
                for b, qp in enumerate(self.qps):
                    buf.a[b, :] = f(qp)
                self._raw.append(buf)
        os_, nb, nbytes = h.load_problem(*[r.a for r in self._raw], opt.bound_relax_factor, opt.bound_push, opt.bound_fac, opt.scaling)
        self.obj_scale[:] = os_
        self.norm_b = nb
        self._c0 = np.array([qp.c0 for qp in self.qps]) * self.obj_scale
        self._con_scale_on_device = True
        self.h2d_bytes = nbytes
        if self.keep_snapshot:
            h.snapshot_inputs()

    def upload(self):
        """Host -> device copy of one solve's inputs (iterate, bounds, rhs, cost, Jacobian/Hessian values)."""
        h, hb = self.h, self._host
        self.h2d_bytes = 0
        for name in ("X", "XL", "XU", "Y", "RHS", "COST"):
            h.vec_set(name, hb[name])
            self.h2d_bytes += hb[name].nbytes
        h.set_values(hb["HV"], hb["JV"])                     # compress_jacobian! / compress_hessian!
        self.h2d_bytes += hb["HV"].nbytes + hb["JV"].nbytes
        if self.keep_snapshot:
            h.snapshot_inputs()

    def start(self):
        """Device-side part of initialize! (solver.jl:144, 162-179); inputs already resident in HBM."""
        opt, h, B = self.opt, self.h, self.B
        if (opt.growth_limit != 1e13 or opt.tiny_pivot != 1e-13) and opt.kkt_system != "NORMAL":   # (the library's defaults)
            h.set_pivot_checks(opt.tiny_pivot, opt.growth_limit)
        self.k = 0
        self.iters[:] = 0
        self.status = [INITIAL] * B
        self.active[:] = 1
        self.trace = []
        self.alpha_p = np.zeros(B)
        self.alpha_d = np.zeros(B)
        for name in ("ZL", "ZU", "JACL", "D"):
            h.vec_fill(name, 0.0)
        self.kkt.initialize()                                # solver.jl:162
        self._init_regularization()                          # solver.jl:163
        self.obj_val = self._eval_callbacks()                # solver.jl:166-169
        self.norm_c = h.vec_norm_inf("F")                    # solver.jl:174
        self.init_starting_point()                           # solver.jl:177
        self.mu[:] = opt.mu_init                             # solver.jl:179

    def _eval_callbacks(self):
        # eval_f / eval_cons / eval_grad_f wrappers in one fused device pass (solver.jl:338-340)
        if self.B == 1:
            return self.h.eval_model(float(self._c0[0]))
        return self.h.eval_model(0.0) + self._c0

    def _init_regularization(self):
        """kernels.jl:380-408."""
        r = self.opt.regularization
        self.del_w = 1.0
        self.del_c = 0.0 if isinstance(r, NoRegularization) else r.delta_d
        self._reg = None if isinstance(r, NoRegularization) else type(r)(**r.__dict__)

    def _update_regularization(self):
        """kernels.jl:386-417."""
        r = self._reg
        if r is None:
            self.del_w, self.del_c = 0.0, 0.0
        elif isinstance(r, FixedRegularization):
            self.del_w, self.del_c = r.delta_p, r.delta_d
        else:
            r.delta_p = max(r.delta_p / 10.0, r.delta_min)
            r.delta_d = min(r.delta_d / 10.0, -r.delta_min)
            self.del_w, self.del_c = r.delta_p, r.delta_d

    # ---- linear_solver.jl ------------------------------------------------------------------------
    def factorize_wrapper(self):
        """MadNLP.factorize_wrapper!: build_kkt! + factorize!."""
        self.kkt.build_kkt()
        return self.kkt.factorize()

    def _bump_regularization(self):
        """linear_solver.jl:14-15 (+ the dual kick described at IPMOptions.dual_regularization_kick)."""
        if self.opt.dual_regularization_kick and self.del_c == 0.0 and self.del_w > 0.0 and self.opt.kkt_system != "NORMAL":
            self.del_c = -self.del_w          # first retry: only the missing dual regularisation (a larger del_w slows the
            return                            # convergence of the dual residual: r_d after a full step is -del_w * dx)
        self.del_w *= 100.0
        self.del_c *= 100.0

    def _garbage_direction(self, norms):
        tol = self.opt.residual_breakdown_tol
        if tol <= 0.0:
            return False
        act = self.active.astype(bool)
        ratio = norms[:, 0] / np.maximum(1.0, norms[:, 1])
        return bool(np.any(ratio[act] > tol))

    def factorize_regularized_system(self, tries=3):
        """/root/reference/src/linear_solver.jl:6-17."""
        for _ in range(tries):
            self.h.set_aug_diagonal_reg(self.del_w, self.del_c)
            self.factorize_wrapper()
            if np.all(self.kkt.is_factorized()[self.active.astype(bool)]):
                break
            self._bump_regularization()

    def solve_system(self):
        """/root/reference/src/linear_solver.jl:19-45 on (d, p)."""
        rc, norms = self.h.solve_system("D", "P")
        self._check_solve_norms(norms)

    def _check_solve_norms(self, norms):
        ratio = norms[:, 0] / np.maximum(1.0, norms[:, 1])
        act = self.active.astype(bool)
        self.last_residual_ratio = ratio
        if np.any(np.isnan(ratio[act])) or (self.opt.check_residual and np.any(ratio[act] > self.opt.tol_linear_solve)):
            raise SolveException()

    # ---- solver.jl:6-125 ---------------------------------------------------------------------------
    def init_starting_point(self):
        h = self.h
        h.set_initial_regularization(self.del_w, self.del_c)        # l.16-18
        self.factorize_wrapper()                                    # l.21
        h.set_initial_primal_rhs()                                  # l.25
        rc, norms = h.solve_system("D", "P")                        # l.26
        if self.opt.dual_regularization_kick or self.opt.residual_breakdown_tol > 0.0:
            # (deviation, see IPMOptions: the reference factorises once here; a breakdown or a garbage solve gets one retry
            #  with the dual kick, otherwise the start is garbage along the dependent rows and the solve never recovers)
            act = self.active.astype(bool)
            if not np.all(self.kkt.is_factorized()[act]) or self._garbage_direction(norms):
                self._bump_regularization()
                h.set_initial_regularization(self.del_w, self.del_c)
                self.factorize_wrapper()
                h.set_initial_primal_rhs()
                rc, norms = h.solve_system("D", "P")
        self._check_solve_norms(norms)
        h.vec_axpy("X", 0, "D", 0, self.n, 1.0)                     # l.28
        h.set_initial_dual_rhs()                                    # l.31
        self.solve_system()                                         # l.32
        h.vec_fill("Y", 0.0)
        h.vec_axpy("Y", 0, "D", self.n, self.m, 1.0)                # l.33  y .= dual(d)
        try:
            h.init_multipliers_and_shift(self.opt.bound_fac)        # l.35-123
        except MadIPMError as e:
            if e.code == _lib.INTERNAL:
                raise AssertionError(str(e))
            raise

    # ---- kernels.jl:226-236, 307-374 -----------------------------------------------------------------
    def update_barrier(self, mu_affine, mu_curr):
        has_inequalities = (self.nlb + self.nub) > 0       # SURVEY note B.1
        if has_inequalities:
            with np.errstate(divide="ignore", invalid="ignore"):
                r = mu_affine / mu_curr
                sigma = np.clip(r * r * r, 1e-6, 10.0)       # Julia lowers ^3 with a literal exponent to x*x*x
        else:
            sigma = np.ones(self.B)
        self.mu = np.maximum(self.opt.mu_min, sigma * mu_curr)

    def get_fraction_to_boundary_step(self, tau):
        r = self.h.ratio_test(tau)
        return np.minimum(r[:, 0], r[:, 1]), np.minimum(r[:, 2], r[:, 3])

    def update_step(self):
        rule = self.opt.step_rule
        if isinstance(rule, ConservativeStep):
            self.alpha_p, self.alpha_d = self.get_fraction_to_boundary_step(rule.tau)
        elif isinstance(rule, AdaptiveStep):
            tau = np.maximum(1 - self.mu, rule.tau_min)
            self.alpha_p, self.alpha_d = self.get_fraction_to_boundary_step(tau)
        elif isinstance(rule, MehrotraAdaptiveStep):
            s = self.h.mehrotra_step(rule.gamma_f)
            self.alpha_p, self.alpha_d = s[:, 0].copy(), s[:, 1].copy()
        else:
            raise TypeError(rule)

    def gondzio_correction_direction(self, mu_curr, max_ncorr):
        """/root/reference/src/solver.jl:200-251 (batch: every instance follows instance-wise accept/reject
        only when batch == 1; for batch > 1 the correction is kept if all active instances accept it)."""
        delta, bmin, bmax, tau = 0.1, 0.1, 10.0, 0.995
        h = self.h
        alpha_p, alpha_d = self.get_fraction_to_boundary_step(tau)
        act = self.active.astype(bool)
        for _ in range(max_ncorr):
            ta_p = np.minimum(alpha_p + delta, 1.0)
            ta_d = np.minimum(alpha_d + delta, 1.0)
            ga = h.complementarity(ta_p, ta_d)[:, 0]
            with np.errstate(divide="ignore", invalid="ignore"):
                mu = (ga / mu_curr) ** 2 * ga
            h.set_extra_correction(ta_p, ta_d, bmin, bmax, mu)
            h.set_correction_rhs(mu)
            h.vec_copy("W2", "D")
            self.solve_system()
            hp, hd = self.get_fraction_to_boundary_step(tau)
            if np.any((hp < 1.005 * alpha_p)[act]) or np.any((hd < 1.005 * alpha_d)[act]):
                h.vec_copy("D", "W2")
                break
            alpha_p, alpha_d = hp, hd
        return alpha_p, alpha_d

    # ---- solver.jl:254-345 -----------------------------------------------------------------------------
    def mpc(self):
        opt, h = self.opt, self.h
        self._obj_pending = False
        self._next_res = None
        while True:
            if self._next_res is not None:                              # left by madipm_phase_iteration: l.259-274 of the new iterate
                res, obj = self._next_res
                self._next_res = None
                self.obj_val = np.where(self.active.astype(bool), obj, self.obj_val)
            elif self._obj_pending:                                       # l.259-274 + the objective of l.338
                res, obj = h.phase_residuals_obj(float(self._c0[0]) if self.B == 1 else 0.0)
                if self.B > 1:
                    obj = obj + self._c0
                self.obj_val = np.where(self.active.astype(bool), obj, self.obj_val)
                self._obj_pending = False
            else:
                res = h.phase_residuals()                               # l.259-274
            self.inf_pr = res[:, 0] / np.maximum(1.0, self.norm_b)
            self.inf_du = res[:, 1] / np.maximum(1.0, self.norm_c)
            self.inf_compl = res[:, 2] / np.maximum(1.0, self.norm_c)
            self.trace.append(dict(k=self.k, obj=(self.obj_val / self.obj_scale).copy(), inf_pr=self.inf_pr.copy(),
                                   inf_du=self.inf_du.copy(), inf_compl=self.inf_compl.copy(), mu=self.mu.copy(),
                                   alpha_p=np.array(self.alpha_p, copy=True), alpha_d=np.array(self.alpha_d, copy=True),
                                   norm_dx=res[:, 3].copy()))
            err = np.maximum(self.inf_pr, np.maximum(self.inf_du, self.inf_compl))
            for b in range(self.B):                                     # l.279-283
                if self.active[b]:
                    if err[b] <= opt.tol:
                        self.status[b] = SOLVE_SUCCEEDED; self.active[b] = 0; self.iters[b] = self.k
                    elif self.k >= opt.max_iter:
                        self.status[b] = MAXIMUM_ITERATIONS_EXCEEDED; self.active[b] = 0; self.iters[b] = self.k
            if not self.active.any():
                return
            self._update_regularization()                               # l.288
            fused = opt.fused_phases
            rule = opt.step_rule
            if (fused and opt.fused_iteration and opt.fused_factor_forward and opt.max_ncorr == 0
                    and isinstance(rule, (ConservativeStep, AdaptiveStep)) and not opt.check_residual):
                # the whole iteration (l.288-342) behind one synchronisation; the device leaves the iterate alone and says so
                # when a factorisation broke down or a residual is not a number: then the phases below redo the iteration
                adaptive = isinstance(rule, AdaptiveStep)
                ok, it = h.phase_iteration(self.del_w, self.del_c, opt.mu_min, rule.tau_min if adaptive else rule.tau, adaptive,
                                           (self.nlb + self.nub) > 0, float(self._c0[0]) if self.B == 1 else 0.0,
                                           None if self.B == 1 else self.active, res_tol=opt.residual_breakdown_tol)
                if ok:
                    self.last_residual_ratio = it[:, 9] / np.maximum(1.0, it[:, 10])
                    self.mu = it[:, 7].copy()
                    self.alpha_p, self.alpha_d = it[:, 20].copy(), it[:, 21].copy()
                    self._next_res = (it[:, 22:26].copy(), it[:, 26] + (self._c0 if self.B > 1 else 0.0))
                    self.k += 1
                    continue
            if fused:
                # first try of factorize_regularized_system without a synchronisation: the breakdown flags come back with
                # the scalars of the predictor phase; a failure (rare) re-enters the reference's retry loop
                if opt.fused_factor_forward:
                    # ... and with the predictor's forward sweep inside the factor launch (three sweeps per iteration)
                    rc, norms, aff = h.phase_factor_predictor(self.del_w, self.del_c)
                else:
                    h.set_aug_diagonal_reg(self.del_w, self.del_c)
                    self.kkt.build_kkt()
                    h.factorize_async()
                    rc, norms, aff = h.phase_predictor()                # l.294-297 affine_direction!
                for _ in range(2):                                      # the retries of linear_solver.jl:8-16
                    if np.all(self.kkt.is_factorized()[self.active.astype(bool)]) and not self._garbage_direction(norms):
                        break
                    self._bump_regularization()
                    self.factorize_regularized_system(tries=1)
                    rc, norms, aff = h.phase_predictor()
                self._check_solve_norms(norms)
            else:
                self.factorize_regularized_system()                     # l.289
                h.set_predictive_rhs()                                  # l.294 affine_direction!
                for _ in range(2):
                    rc, norms = self.h.solve_system("D", "P")
                    if not self._garbage_direction(norms):
                        break
                    self._bump_regularization()
                    self.factorize_regularized_system(tries=1)
                self._check_solve_norms(norms)
                aff = h.phase_affine()                                  # l.295-297
            mu_affine, mu_curr = aff[:, 2], aff[:, 3]
            self.update_barrier(mu_affine, mu_curr)                     # l.302
            rule = opt.step_rule
            if fused and opt.max_ncorr == 0 and isinstance(rule, (ConservativeStep, AdaptiveStep)):
                tau = rule.tau if isinstance(rule, ConservativeStep) else np.maximum(1 - self.mu, rule.tau_min)
                rc, norms, r = h.phase_corrector(self.mu, tau)          # l.307-311 + l.329
                self._check_solve_norms(norms)
                self.alpha_p, self.alpha_d = np.minimum(r[:, 0], r[:, 1]), np.minimum(r[:, 2], r[:, 3])
            else:
                h.set_correction_rhs(self.mu)                           # l.307-311
                self.solve_system()
                if opt.max_ncorr > 0:                                   # l.316-324
                    self.gondzio_correction_direction(mu_curr, opt.max_ncorr)
                self.update_step()                                      # l.329
            h.update_iterate(self.alpha_p, self.alpha_d, self.mu,       # l.332-335 and 342
                             None if self.B == 1 else self.active)
            if fused:
                h.eval_model(want_obj=False)                            # l.338-340; the objective is read with the next residuals
                self._obj_pending = True
            else:
                obj = self._eval_callbacks()
                act = self.active.astype(bool)
                self.obj_val = np.where(act, obj, self.obj_val)
            self.k += 1

    # ---- solver.jl:347-403 -------------------------------------------------------------------------------
    def solve(self, resident=False):
        """solve!(solver).  resident=True skips the host preparation and the H2D copies (the inputs of a
        previous `upload` are reused), which is what bench.py times as the device-resident figure."""
        t0 = time.perf_counter()
        self.init_time = 0.0
        try:
            if resident:
                self.upload_resident()
            elif self.opt.device_prepare:
                self.load_device()
            else:
                self.prepare_host()
                self.upload()
            self.start()
            self.init_time = time.perf_counter() - t0
            self.mpc()
        except SolveException:
            self._fail(ERROR_IN_STEP_COMPUTATION)
        except AssertionError:
            self._fail(INTERNAL_ERROR)
        except MadIPMError as e:
            self.last_error_message = str(e)
            if e.code == _lib.CUDA_ERROR:
                raise
            if e.code == _lib.INTERNAL:
                import sys
                sys.stderr.write(f"[madipm] {e}\n")
                sys.stderr.flush()
                if "still busy" in str(e) or "wedged" in str(e):
                    raise                  # the stream holds a kernel that never returns: nothing more can be read from this handle
            self._fail(ERROR_IN_STEP_COMPUTATION if e.code in (_lib.NONFINITE, _lib.FACTOR_BREAKDOWN) else INTERNAL_ERROR)
        if getattr(self, "_res", None) is None:              # page-locked landing buffers for the results
            self._res = {v: _lib.PinnedArray((self.B, self.h.vec_len(v))) for v in ("X", "Y", "CON_SCALE")}
        x = self.h.vec_get("X", out=self._res["X"].a).reshape(self.B, -1)
        y = self.h.vec_get("Y", out=self._res["Y"].a).reshape(self.B, -1)
        self.d2h_bytes = x.nbytes + y.nbytes
        if getattr(self, "_con_scale_on_device", False):
            self.con_scale = self.h.vec_get("CON_SCALE", out=self._res["CON_SCALE"].a).reshape(self.B, -1).copy()
            self.d2h_bytes += self.con_scale.nbytes
        self.total_time = time.perf_counter() - t0
        sol = x[:, :self.nx].copy()
        if self._fixed.size:                                   # the parameters back in their places
            full = np.empty((self.B, self.nx_full))
            full[:, self._free] = sol
            full[:, self._fixed] = self._x_fixed
            sol = full
        return ExecutionStats(
            status=list(self.status), iter=self.iters.copy(), objective=self.obj_val / self.obj_scale,
            solution=sol, multipliers=y * self.con_scale / self.obj_scale[:, None],
            inf_pr=self.inf_pr.copy(), inf_du=self.inf_du.copy(), inf_compl=self.inf_compl.copy(),
            total_time=self.total_time, init_time=self.init_time)

    def upload_resident(self):
        """Inputs of the last `upload` are restored from their device-side snapshot (no host traffic)."""
        self.h.restore_inputs()

    def _fail(self, status):
        for b in range(self.B):
            if self.active[b]:
                self.status[b] = status
                self.iters[b] = self.k
        self.active[:] = 0

    def close(self):
        self.h.close()
