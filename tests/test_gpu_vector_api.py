"""Sub-range / index-view entry points behind the Julia vector type (julia/MadIPMB200.jl): the lines of
/root/reference/src/solver.jl that work on solver vectors directly (axpy! :28, :332-333; `.=` :16-18, :33; `.+=` on the
zl_r / zu_r views :334-335; adjust_boundary! :342) against numpy, and against the fused madipm_update_iterate."""
import numpy as np
import pytest

from madqp_jl_b200 import problems as P
from madqp_jl_b200.mpc import FixedRegularization, MPCSolver

pytestmark = pytest.mark.gpu


def _solver(batch=1):
    qps = [P.small_qp(instance=i) for i in range(batch)] if batch > 1 else [P.inequality_qp(seed=3, n=60, m=40, n_ineq=15)]
    s = MPCSolver(qps if batch > 1 else qps[0], regularization=FixedRegularization(1e-8, -1e-8))
    s.initialize()
    return s


@pytest.mark.parametrize("batch", [1, 3])
def test_range_get_set_fill_copy_norm(batch):
    s = _solver(batch)
    h = s.h
    L = h.vec_len("D")
    rng = np.random.default_rng(0)
    full = rng.standard_normal((batch, L))
    h.vec_set("D", full)
    off, ln = 7, s.n + 5
    np.testing.assert_array_equal(np.reshape(h.vec_get_range("D", off, ln), (batch, ln)), full[:, off:off + ln])
    piece = rng.standard_normal((batch, ln))
    h.vec_set_range("D", off, piece)
    full[:, off:off + ln] = piece
    np.testing.assert_array_equal(np.reshape(h.vec_get("D"), (batch, L)), full)
    h.vec_fill_range("D", 2, 3, -4.5)
    full[:, 2:5] = -4.5
    h.vec_copy_range("P", s.n, "D", 0, s.m)                          # solver.y .= dual(d) style copies between ranges
    p = np.reshape(h.vec_get("P"), (batch, L))
    np.testing.assert_array_equal(p[:, s.n:s.n + s.m], full[:, :s.m])
    np.testing.assert_array_equal(h.vec_norm_inf_range("D", off, ln), np.max(np.abs(full[:, off:off + ln]), axis=1))
    np.testing.assert_array_equal(h.vec_norm_inf_range("D", 0, 0), np.zeros(batch))   # empty range
    with pytest.raises(Exception):
        h.vec_get_range("D", L - 1, 2)
    s.close()


def test_unfused_iterate_update_equals_the_fused_entry_point():
    """solver.jl:332-335 + :342 through the separate vector operations == madipm_update_iterate (bit for bit)."""
    a, b = _solver(), _solver()
    rng = np.random.default_rng(1)
    L = a.h.vec_len("D")
    d = rng.standard_normal(L) * 1e-3
    ap, ad, mu = 0.7, 0.4, 1e16         # a huge mu (eps * mu ~ 2) makes adjust_boundary! move bounds, so that branch is exercised
    for s in (a, b):
        s.h.vec_set("D", d)
    n, m, nlb = a.n, a.m, a.nlb
    a.h.update_iterate(np.array([ap]), np.array([ad]), np.array([mu]), None)
    h = b.h
    h.vec_axpy("X", 0, "D", 0, n, ap)                                # axpy!(alpha_p, primal(d), primal(x))
    h.vec_axpy("Y", 0, "D", n, m, ad)                                # axpy!(alpha_d, dual(d), y)
    h.vec_axpy_at("ZL", 0, "D", n + m, ad)                           # zl_r .+= alpha_d .* dual_lb(d)
    h.vec_axpy_at("ZU", 1, "D", n + m + nlb, ad)                     # zu_r .+= alpha_d .* dual_ub(d)
    h.adjust_boundary(mu)
    moved = 0
    for v in ("X", "Y", "ZL", "ZU", "XL", "XU"):
        np.testing.assert_array_equal(a.h.vec_get(v), b.h.vec_get(v), err_msg=v)
    moved = np.sum(a.h.vec_get("XL") != _solver().h.vec_get("XL"))
    assert moved > 0
    a.close(); b.close()
