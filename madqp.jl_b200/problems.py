"""Seeded synthetic LP/QP instances of the shapes named in BASELINE.json.

The reference ships no generator (its benchmarks read netlib / Maros-Meszaros /
MIPLIB files, /root/reference/scripts/benchmarks_cpu.jl:66-76), so these are
builder-defined; the rules follow SURVEY.md section 8(d):

* ``numpy.random.default_rng(seed)``; no duplicate (i, j) entries in A (the
  reference's normal-equations assembly overwrites on duplicates,
  /root/reference/src/utils.jl:280-283);
* nonzeros U(-1, 1) with |v| >= 0.1;
* feasible and bounded by construction: choose a primal point x*, a random
  active set, multipliers y*, z*, then set ``b = A x*`` and
  ``c = -Q x* - A'y* + zl* - zu*`` so that (x*, y*, z*) is a KKT point of

      min c0 + c'x + 1/2 x'Qx   s.t.  lcon <= A x <= ucon,  lvar <= x <= uvar

  which is the QuadraticModels form the reference consumes
  (/root/reference/test/runtests.jl:24-55).

Indices are 0-based here; the C-ABI layer shifts to the 1-based Int32 triplets
a Julia caller would pass.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass, field

import numpy as np

INF = float("inf")


@dataclass
class QPData:
    name: str
    nvar: int
    ncon: int
    c: np.ndarray
    c0: float
    # lower triangle of the Hessian, COO
    H_rows: np.ndarray
    H_cols: np.ndarray
    H_vals: np.ndarray
    # constraint matrix, COO
    A_rows: np.ndarray
    A_cols: np.ndarray
    A_vals: np.ndarray
    lcon: np.ndarray
    ucon: np.ndarray
    lvar: np.ndarray
    uvar: np.ndarray
    x0: np.ndarray
    y0: np.ndarray
    meta: dict = field(default_factory=dict)

    @property
    def nnzj(self) -> int:
        return int(self.A_rows.size)

    @property
    def nnzh(self) -> int:
        return int(self.H_rows.size)


def _values(rng, size):
    v = rng.uniform(-1.0, 1.0, size)
    small = np.abs(v) < 0.1
    v[small] = np.where(v[small] < 0, -0.1, 0.1) + 0.0
    return v


def _window_rows(rng, m, n, per_row, half_width, stride):
    """``per_row`` distinct columns for every row i drawn from the window
    [stride*i - W, stride*i + W] clipped to [0, n).  Returns (rows, cols) with
    row-major order, columns ascending inside a row."""
    width = 2 * half_width + 1
    centre = (np.arange(m, dtype=np.int64) * stride)[:, None]
    cand = centre + np.arange(-half_width, half_width + 1, dtype=np.int64)[None, :]
    key = rng.random((m, width))
    key[(cand < 0) | (cand >= n)] = 2.0  # never selected while valid ones remain
    pick = np.argpartition(key, per_row - 1, axis=1)[:, :per_row]
    cols = np.take_along_axis(cand, pick, axis=1)
    cols.sort(axis=1)
    rows = np.repeat(np.arange(m, dtype=np.int64), per_row)
    return rows, cols.reshape(-1)


def _finish(name, rng, n, m, A_rows, A_cols, A_vals, H_rows, H_cols, H_vals,
            upper_fraction, meta, bound_rng=None):
    """Pick the KKT point and derive b, c, bounds (SURVEY 8d common rules).
    ``bound_rng`` fixes WHICH variables are boxed independently of the values (batch members
    must share the bound structure as well as the sparsity pattern)."""
    lvar = np.zeros(n)
    uvar = np.full(n, INF)
    has_ub = (bound_rng if bound_rng is not None else rng).random(n) < upper_fraction
    uvar[has_ub] = 10.0

    x_star = rng.uniform(0.5, 1.5, n)
    active = rng.random(n) < 0.3
    at_upper = active & has_ub & (rng.random(n) < 0.5)
    at_lower = active & ~at_upper
    x_star[at_lower] = 0.0
    x_star[at_upper] = 10.0
    zl = np.zeros(n)
    zu = np.zeros(n)
    zl[at_lower] = rng.uniform(0.1, 1.0, int(at_lower.sum()))
    zu[at_upper] = rng.uniform(0.1, 1.0, int(at_upper.sum()))
    y_star = rng.standard_normal(m)

    b = np.zeros(m)
    np.add.at(b, A_rows, A_vals * x_star[A_cols])
    aty = np.zeros(n)
    np.add.at(aty, A_cols, A_vals * y_star[A_rows])
    qx = np.zeros(n)
    np.add.at(qx, H_rows, H_vals * x_star[H_cols])
    off = H_rows != H_cols
    np.add.at(qx, H_cols[off], H_vals[off] * x_star[H_rows[off]])
    # stationarity in the reference's sign convention: f + A'y - zl + zu = 0
    # with f = Qx + c   (/root/reference/src/kernels.jl:36)
    c = -qx - aty + zl - zu
    obj_star = float(c @ x_star + 0.5 * (x_star @ qx))

    meta = dict(meta)
    meta.update(obj_star=obj_star, n_active=int(active.sum()))
    return QPData(
        name=name, nvar=n, ncon=m, c=c, c0=0.0,
        H_rows=H_rows.astype(np.int64), H_cols=H_cols.astype(np.int64), H_vals=H_vals,
        A_rows=A_rows.astype(np.int64), A_cols=A_cols.astype(np.int64), A_vals=A_vals,
        lcon=b.copy(), ucon=b.copy(), lvar=lvar, uvar=uvar,
        x0=np.zeros(n), y0=np.zeros(m), meta=meta,
    )


def _banded_q(rng, n, bandwidth):
    """SPD banded Q by diagonal dominance, returned as lower-triangle COO."""
    rows = [np.arange(n)]
    cols = [np.arange(n)]
    vals = [2.0 + rng.random(n)]
    off = [-0.5, -0.25, -0.125]
    for k in range(1, bandwidth + 1):
        rows.append(np.arange(k, n))
        cols.append(np.arange(0, n - k))
        vals.append(np.full(n - k, off[k - 1]))
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def _empty_q():
    e = np.zeros(0, dtype=np.int64)
    return e, e.copy(), np.zeros(0)


def maros_meszaros_like(seed=0, n=1000, m=500, nnz=5000):
    """C1: n=1k, m=500, ~1% dense A at uniformly random distinct positions with
    every row non-empty; tridiagonal SPD Q; half the variables boxed."""
    rng = np.random.default_rng(seed)
    first = rng.integers(0, n, m)  # one entry per row guarantees non-empty rows
    taken = set((int(i) * n + int(j)) for i, j in zip(range(m), first))
    need = nnz - m
    while len(taken) < nnz:
        extra = rng.integers(0, m * n, 2 * need)
        for e in extra:
            if len(taken) >= nnz:
                break
            taken.add(int(e))
    lin = np.array(sorted(taken), dtype=np.int64)
    A_rows, A_cols = lin // n, lin % n
    A_vals = _values(rng, lin.size)
    H = _banded_q(rng, n, 1)
    return _finish(f"mm_like_n{n}_m{m}_s{seed}", rng, n, m, A_rows, A_cols, A_vals, *H,
                   upper_fraction=0.5, meta=dict(config="C1", seed=seed))


def sparse_lp(seed=0, n=200_000, m=100_000, per_row=10, half_width=64, long_range=1,
              long_range_window=1024):
    """C2: LP, ``per_row`` nonzeros per row: ``per_row - long_range`` in the
    window [2i-W, 2i+W] plus ``long_range`` longer-range columns drawn from
    [2i-R, 2i+R], R = ``long_range_window`` (``None`` = uniform over all
    columns).  SURVEY 8d proposed uniform long-range columns; measured at 1/10
    scale that makes A A' an expander whose factor ends in a dense front of 0.69 m
    columns (2.4e9 nnz(L), 1e14 flops at full size), which no CPU baseline can
    factor, so the default keeps the long-range coupling bounded."""
    rng = np.random.default_rng(seed)
    stride = max(1, n // m)
    local = per_row - long_range
    rows, cols = _window_rows(rng, m, n, local, half_width, stride)
    if long_range:
        cols2d = cols.reshape(m, local)
        def draw(shape_rows):
            if long_range_window is None:
                return rng.integers(0, n, shape_rows)
            c = stride * np.arange(m, dtype=np.int64)[:, None] + rng.integers(
                -long_range_window, long_range_window + 1, (m, long_range))
            return np.clip(c, 0, n - 1)
        far = draw((m, long_range))
        # re-draw collisions with the window picks (no duplicate (i, j))
        for _ in range(8):
            clash = (far[:, :, None] == cols2d[:, None, :]).any(axis=2)
            if long_range > 1:
                srt = np.sort(far, axis=1)
                dup = np.zeros_like(clash)
                dup[:, 1:] = srt[:, 1:] == srt[:, :-1]
                far = srt
                clash = (far[:, :, None] == cols2d[:, None, :]).any(axis=2) | dup
            if not clash.any():
                break
            far[clash] = draw((m, long_range))[clash]
        cols2d = np.concatenate([cols2d, far], axis=1)
        cols2d.sort(axis=1)
        cols = cols2d.reshape(-1)
        rows = np.repeat(np.arange(m, dtype=np.int64), per_row)
    A_vals = _values(rng, rows.size)
    return _finish(f"sparse_lp_n{n}_m{m}_s{seed}", rng, n, m, rows, cols, A_vals, *_empty_q(),
                   upper_fraction=0.0,
                   meta=dict(config="C2", seed=seed, half_width=half_width, long_range=long_range,
                             long_range_window=long_range_window))


def banded_qp(seed=0, n=500_000, m=250_000, per_row=4, half_width=32, q_bandwidth=2):
    """C3: convex QP, 4 nonzeros per row of A in the window [2i-W, 2i+W], Q
    pentadiagonal SPD, half the variables boxed to [0, 10]."""
    rng = np.random.default_rng(seed)
    stride = max(1, n // m)
    rows, cols = _window_rows(rng, m, n, per_row, half_width, stride)
    A_vals = _values(rng, rows.size)
    H = _banded_q(rng, n, q_bandwidth)
    return _finish(f"banded_qp_n{n}_m{m}_s{seed}", rng, n, m, rows, cols, A_vals, *H,
                   upper_fraction=0.5,
                   meta=dict(config="C3", seed=seed, half_width=half_width, q_bandwidth=q_bandwidth))


def transportation_lp(seed=0, n_src=25_000, n_dst=25_000, n_arcs=2_000_000):
    """C4: sparse transportation LP.  Arc k joins source s_k and sink d_k; its
    column has +1 in row s_k and +1 in row n_src + d_k.  Costs U(1, 10) are
    replaced by the KKT-consistent c of ``_finish`` shifted to be positive
    where inactive; supplies/demands come from b = A x*."""
    rng = np.random.default_rng(seed)
    total = n_src * n_dst
    if n_arcs > total:
        raise ValueError("more arcs than source-sink pairs")
    pick = np.unique(rng.integers(0, total, int(n_arcs * 1.05) + 16))
    while pick.size < n_arcs:
        pick = np.unique(np.concatenate([pick, rng.integers(0, total, n_arcs)]))
    pick = rng.permutation(pick)[:n_arcs]
    pick.sort()
    src, dst = pick // n_dst, pick % n_dst
    n, m = n_arcs, n_src + n_dst
    k = np.arange(n, dtype=np.int64)
    A_rows = np.concatenate([src, n_src + dst])
    A_cols = np.concatenate([k, k])
    A_vals = np.ones(2 * n)
    return _finish(f"transport_{n_src}x{n_dst}_a{n_arcs}_s{seed}", rng, n, m, A_rows, A_cols, A_vals,
                   *_empty_q(), upper_fraction=0.0,
                   meta=dict(config="C4", seed=seed, n_src=n_src, n_dst=n_dst))


@functools.lru_cache(maxsize=8)
def small_qp_pattern(seed=0, n=200, m=100, nnz=2000):
    """Pattern shared by the C5 batch (10% dense A, tridiagonal Q)."""
    rng = np.random.default_rng(seed)
    first = rng.integers(0, n, m)
    lin = set(int(i) * n + int(j) for i, j in zip(range(m), first))
    while len(lin) < nnz:
        for e in rng.integers(0, m * n, nnz):
            if len(lin) >= nnz:
                break
            lin.add(int(e))
    lin = np.array(sorted(lin), dtype=np.int64)
    return lin // n, lin % n


def small_qp(instance=0, pattern_seed=0, n=200, m=100, nnz=2000):
    """One member of the C5 batch: pattern fixed by ``pattern_seed``, values and
    KKT point seeded by the instance id."""
    A_rows, A_cols = small_qp_pattern(pattern_seed, n, m, nnz)
    rng = np.random.default_rng(1_000_003 + instance)
    A_vals = _values(rng, A_rows.size)
    H = _banded_q(rng, n, 1)
    return _finish(f"small_qp_{instance}", rng, n, m, A_rows, A_cols, A_vals, *H,
                   upper_fraction=0.5, meta=dict(config="C5", instance=instance),
                   bound_rng=np.random.default_rng(7_000_003 + pattern_seed))


def simple_lp():
    """The only fully specified instance in the reference:
    min x1 + x2  s.t.  x1 + x2 = 1, x >= 0, x0 = (1, 1); objective 1.0
    (/root/reference/test/runtests.jl:24-55)."""
    e = np.zeros(0, dtype=np.int64)
    return QPData(
        name="simpleLP", nvar=2, ncon=1, c=np.ones(2), c0=0.0,
        H_rows=e, H_cols=e.copy(), H_vals=np.zeros(0),
        A_rows=np.array([0, 0]), A_cols=np.array([0, 1]), A_vals=np.array([1.0, 1.0]),
        lcon=np.array([1.0]), ucon=np.array([1.0]),
        lvar=np.zeros(2), uvar=np.full(2, INF), x0=np.ones(2), y0=np.zeros(1),
        meta=dict(obj_star=1.0),
    )


def inequality_qp(seed=0, n=40, m=25, n_ineq=10):
    """Small QP with ``n_ineq`` range/inequality rows so the slack columns of the
    KKT structure (/root/reference/src/KKT/normalkkt.jl:74-77) are exercised."""
    rng = np.random.default_rng(seed)
    dense = rng.random((m, n)) < 0.25
    dense[np.arange(m), rng.integers(0, n, m)] = True
    A_rows, A_cols = np.nonzero(dense)
    A_vals = _values(rng, A_rows.size)
    H = _banded_q(rng, n, 1)
    qp = _finish(f"ineq_qp_{seed}", rng, n, m, A_rows, A_cols, A_vals, *H,
                 upper_fraction=0.5, meta=dict(seed=seed))
    idx = rng.permutation(m)[:n_ineq]
    qp.lcon[idx] -= rng.uniform(0.5, 1.0, n_ineq)   # Ax* strictly inside the range
    half = idx[: n_ineq // 2]
    qp.ucon[idx] += rng.uniform(0.5, 1.0, n_ineq)
    qp.ucon[half] = INF                              # one-sided rows
    return qp


CONFIGS = {
    "C1": maros_meszaros_like,
    "C2": sparse_lp,
    "C3": banded_qp,
    "C4": transportation_lp,
}
