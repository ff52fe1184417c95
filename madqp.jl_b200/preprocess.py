"""Preprocessing of scripts/common.jl through the C ABI, on the device:
  * scale_qp (/root/reference/scripts/common.jl:46-100): Ruiz equilibration of the constraint matrix and the rescaled
    model (`madipm_scale_qp`, csrc/preprocess.cu);
  * standard_form_qp (/root/reference/scripts/common.jl:156-288): slack variables for the inequality rows, range bounds
    moved into equality rows (`madipm_standard_form_qp`, same file).
There is no CPU fallback: without a GPU both calls raise.  `standard_form_qp_host` is the same reformulation in numpy for
callers that hold the model on the host anyway (the reference's own is host code); it is never chosen automatically.
The loop-for-loop restatements the tests compare against are `oracle.oracle.scale_qp` / `oracle.oracle.standard_form_qp`."""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import _lib


def scale_qp(qp, max_iter=None, tol=None, device=0):
    """Returns (scaled_qp, Dr, Dc, info): As = A ./ (Dr_i Dc_j), Hs = H ./ (Dc_i Dc_j), c ./ Dc, lvar/uvar/x0 .* Dc,
    lcon/ucon/y0 ./ Dr (common.jl:70-97).  `info` = dict(sweeps, err, launches)."""
    lib = _lib.load()
    opt = _lib.ScaleOptions()
    lib.madipm_default_scale_options(C.byref(opt))
    opt.device = device
    if max_iter is not None:
        opt.max_iter = int(max_iter)
    if tol is not None:
        opt.tol = float(tol)
    f = lambda a: np.array(a, dtype=np.float64, copy=True)           # the entry point rescales in place
    i32 = lambda a: np.ascontiguousarray(np.asarray(a) + 1, dtype=np.int32)
    jI, jJ, jV = i32(qp.A_rows), i32(qp.A_cols), f(qp.A_vals)
    hI, hJ, hV = i32(qp.H_rows), i32(qp.H_cols), f(qp.H_vals)
    c, lvar, uvar, x0 = f(qp.c), f(qp.lvar), f(qp.uvar), f(qp.x0)
    lcon, ucon, y0 = f(qp.lcon), f(qp.ucon), f(qp.y0)
    Dr, Dc = np.empty(qp.ncon), np.empty(qp.nvar)
    sweeps, err = C.c_int32(0), C.c_double(0.0)
    rc = lib.madipm_scale_qp(C.byref(opt), qp.nvar, qp.ncon, jV.size, _lib._p32(jI), _lib._p32(jJ), _lib._pd(jV),
                             hV.size, _lib._p32(hI), _lib._p32(hJ), _lib._pd(hV), _lib._pd(c), _lib._pd(lvar),
                             _lib._pd(uvar), _lib._pd(lcon), _lib._pd(ucon), _lib._pd(x0), _lib._pd(y0), _lib._pd(Dr),
                             _lib._pd(Dc), C.byref(sweeps), C.byref(err))
    if rc != _lib.OK:
        raise _lib.MadIPMError(rc, lib.madipm_scale_last_error().decode())
    out = dataclasses.replace(qp, name=qp.name + "-ruiz", c=c, H_vals=hV, A_vals=jV, lcon=lcon, ucon=ucon, lvar=lvar,
                              uvar=uvar, x0=x0, y0=y0, meta=dict(qp.meta, ruiz=True))
    return out, Dr, Dc, dict(sweeps=sweeps.value, err=err.value, launches=int(opt.launches))


def standard_form_qp(qp, device=0, info=None):
    """standard_form_qp (/root/reference/scripts/common.jl:156-288) on the device (`madipm_standard_form_qp`): the model's
    bounds and COO triplets go up, the bounds are classified, one prefix sum ranks the inequality rows and the range
    bounds, the appended entries / new bounds are written at the ranked positions, and the reformulated model comes
    back -- entry for entry the reference's ([A | slack | range pairs], common.jl:212-233).  Raises without a GPU.
    `info` (a dict) receives ns, nw and the number of kernels launched."""
    lib = _lib.load()
    n, m = int(qp.nvar), int(qp.ncon)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    i32 = lambda a: np.ascontiguousarray(np.asarray(a) + 1, dtype=np.int32)
    jI, jJ, jV = i32(qp.A_rows), i32(qp.A_cols), f(qp.A_vals)
    c, lvar, uvar, x0 = f(qp.c), f(qp.lvar), f(qp.uvar), f(qp.x0)
    lcon, ucon, y0 = f(qp.lcon), f(qp.ucon), f(qp.y0)
    nnzj = int(jV.size)
    cap_var, cap_con, cap_nz = 2 * (n + m), n + 2 * m, nnzj + m + 2 * (n + m)
    oI, oJ, oV = np.empty(cap_nz, dtype=np.int32), np.empty(cap_nz, dtype=np.int32), np.empty(cap_nz)
    oc, olv, ouv, ox0 = (np.empty(cap_var) for _ in range(4))
    olc, ouc, oy0 = (np.empty(cap_con) for _ in range(3))
    ns, nw, launches = C.c_int32(0), C.c_int32(0), C.c_int64(0)
    rc = lib.madipm_standard_form_qp(device, n, m, nnzj, _lib._p32(jI), _lib._p32(jJ), _lib._pd(jV), _lib._pd(c), _lib._pd(lvar),
                                     _lib._pd(uvar), _lib._pd(lcon), _lib._pd(ucon), _lib._pd(x0), _lib._pd(y0), C.byref(ns),
                                     C.byref(nw), _lib._p32(oI), _lib._p32(oJ), _lib._pd(oV), _lib._pd(oc), _lib._pd(olv),
                                     _lib._pd(ouv), _lib._pd(ox0), _lib._pd(olc), _lib._pd(ouc), _lib._pd(oy0), C.byref(launches))
    if rc != _lib.OK:
        raise _lib.MadIPMError(rc, lib.madipm_scale_last_error().decode())
    ns, nw = ns.value, nw.value
    nvar, ncon, nz = n + ns + nw, m + nw, nnzj + ns + 2 * nw
    if info is not None:
        info.update(ns=ns, nw=nw, launches=int(launches.value))
    return dataclasses.replace(
        qp, name=qp.name + "-std", nvar=nvar, ncon=ncon, c=oc[:nvar].copy(),
        A_rows=(oI[:nz] - 1).astype(np.asarray(qp.A_rows).dtype), A_cols=(oJ[:nz] - 1).astype(np.asarray(qp.A_cols).dtype),
        A_vals=oV[:nz].copy(), lcon=olc[:ncon].copy(), ucon=ouc[:ncon].copy(), lvar=olv[:nvar].copy(), uvar=ouv[:nvar].copy(),
        x0=ox0[:nvar].copy(), y0=oy0[:ncon].copy(), meta=dict(qp.meta, standard_form=True))


def standard_form_qp_host(qp):
    """standard_form_qp (/root/reference/scripts/common.jl:129-288) in numpy, host side like the reference's: slack variables
    s = A x for the inequality rows (columns n+1..n+ns, entries -1), and every RANGE bound of x or s moved into an
    equality row  x_i + w_k = xu_i  with a new nonnegative variable w_k (columns n+ns+1..).  Index and copy work only,
    no arithmetic: the result is the reference's model entry for entry (COO order [A | slack | range pairs]).
    Variables with only an upper bound keep it; fixed variables stay fixed (common.jl:254-257)."""
    n, m = qp.nvar, qp.ncon
    lvar, uvar = np.asarray(qp.lvar, dtype=np.float64), np.asarray(qp.uvar, dtype=np.float64)
    lcon, ucon = np.asarray(qp.lcon, dtype=np.float64), np.asarray(qp.ucon, dtype=np.float64)
    ind_ineq = np.flatnonzero(lcon < ucon)                                         # l.164-168
    ns = ind_ineq.size
    fixed = np.flatnonzero(lvar == uvar)                                           # l.177-179
    rng_x = np.flatnonzero((lvar != uvar) & (-np.inf < lvar) & (lvar < uvar) & (uvar < np.inf))   # l.180-183
    li, ui = lcon[ind_ineq], ucon[ind_ineq]
    rng_s = np.flatnonzero((-np.inf < li) & (li < ui) & (ui < np.inf))             # l.191-195
    ind_rng = np.concatenate([rng_x, n + rng_s])                                   # 0-based columns of [x | s]
    xu = np.concatenate([uvar[rng_x], ui[rng_s]])
    nw = ind_rng.size
    nvar, ncon = n + ns + nw, m + nw
    k = np.arange(nw)
    A_rows = np.concatenate([qp.A_rows, ind_ineq, np.repeat(m + k, 2)])            # l.212-233
    A_cols = np.concatenate([qp.A_cols, n + np.arange(ns), np.stack([ind_rng, n + ns + k], axis=1).ravel()])
    A_vals = np.concatenate([np.asarray(qp.A_vals, dtype=np.float64), np.full(ns, -1.0), np.ones(2 * nw)])
    lcon_, ucon_ = np.zeros(ncon), np.zeros(ncon)                                  # l.236-253
    eq = ~(lcon < ucon)
    lcon_[:m][eq], ucon_[:m][eq] = lcon[eq], ucon[eq]
    lcon_[m:], ucon_[m:] = xu, xu
    lvar_ = np.concatenate([lvar, li, np.zeros(nw)])                               # l.255-260
    uvar_ = np.concatenate([uvar, ui, np.full(nw, np.inf)])
    uvar_[ind_rng] = np.inf
    uvar_[fixed] = uvar[fixed]
    z = np.zeros(ns + nw)
    return dataclasses.replace(
        qp, name=qp.name + "-std", nvar=nvar, ncon=ncon, c=np.concatenate([qp.c, z]), A_rows=A_rows.astype(qp.A_rows.dtype),
        A_cols=A_cols.astype(qp.A_cols.dtype), A_vals=A_vals, lcon=lcon_, ucon=ucon_, lvar=lvar_, uvar=uvar_,
        x0=np.concatenate([qp.x0, z]), y0=np.concatenate([qp.y0, np.zeros(nw)]), meta=dict(qp.meta, standard_form=True))


def make_parameter(qp):
    """fixed_variable_treatment = MakeParameter (/root/reference/src/utils.jl:81; MadNLP's SparseCallback, un-vendored):
    variables with lvar == uvar are parameters, not unknowns.  Host-side index work before the hot path, like the
    reference's: the equivalent model in the free variables (c_free += H[free, fixed] x_f, c0 += c_f' x_f + x_f' H_ff x_f / 2,
    lcon / ucon -= A[:, fixed] x_f, the fixed rows / columns dropped from the COO triplets in their order).
    Returns (reduced model, fixed, free, x_fixed) with 0-based index arrays; the model itself when nothing is fixed."""
    n = qp.nvar
    lvar, uvar = np.asarray(qp.lvar, dtype=np.float64), np.asarray(qp.uvar, dtype=np.float64)
    is_fixed = lvar == uvar
    fixed, free = np.flatnonzero(is_fixed), np.flatnonzero(~is_fixed)
    if fixed.size == 0:
        return qp, fixed, free, np.zeros(0)
    xfull = np.where(is_fixed, lvar, 0.0)                         # x_fixed scattered, zero elsewhere
    new_index = np.cumsum(~is_fixed) - 1
    Hr, Hc, Hv = np.asarray(qp.H_rows), np.asarray(qp.H_cols), np.asarray(qp.H_vals, dtype=np.float64)
    Ar, Ac, Av = np.asarray(qp.A_rows), np.asarray(qp.A_cols), np.asarray(qp.A_vals, dtype=np.float64)
    c = np.asarray(qp.c, dtype=np.float64).copy()
    c0 = float(qp.c0) + float(np.dot(c[fixed], lvar[fixed]))
    rf, cf = is_fixed[Hr], is_fixed[Hc]
    both = rf & cf
    c0 += float(np.sum(np.where(Hr[both] == Hc[both], 0.5, 1.0) * Hv[both] * xfull[Hr[both]] * xfull[Hc[both]]))
    only_r, only_c = rf & ~cf, cf & ~rf                           # lower triangle of a symmetric matrix: both orientations
    np.add.at(c, Hc[only_r], Hv[only_r] * xfull[Hr[only_r]])
    np.add.at(c, Hr[only_c], Hv[only_c] * xfull[Hc[only_c]])
    keep_h = ~rf & ~cf
    af = is_fixed[Ac]
    shift = np.bincount(Ar[af], weights=Av[af] * xfull[Ac[af]], minlength=qp.ncon)[:qp.ncon] if qp.ncon else np.zeros(0)
    red = dataclasses.replace(
        qp, name=qp.name + "-free", nvar=int(free.size), c=c[free], c0=c0,
        H_rows=new_index[Hr[keep_h]].astype(Hr.dtype), H_cols=new_index[Hc[keep_h]].astype(Hc.dtype), H_vals=Hv[keep_h],
        A_rows=Ar[~af], A_cols=new_index[Ac[~af]].astype(Ac.dtype), A_vals=Av[~af],
        lcon=np.asarray(qp.lcon, dtype=np.float64) - shift, ucon=np.asarray(qp.ucon, dtype=np.float64) - shift,
        lvar=lvar[free], uvar=uvar[free], x0=np.asarray(qp.x0, dtype=np.float64)[free],
        meta=dict(qp.meta, make_parameter=int(fixed.size)))
    return red, fixed, free, lvar[fixed]
