"""scale_qp (/root/reference/scripts/common.jl:46-100) through the C ABI: Ruiz equilibration of the constraint matrix
on the device (`madipm_scale_qp`, csrc/preprocess.cu) and the rescaled model.  There is no CPU fallback: without a
GPU the call raises.  The numpy restatement used by the tests is `oracle.oracle.scale_qp`."""
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
