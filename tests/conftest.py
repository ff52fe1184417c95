import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_definitions.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything native is built once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()


def analyze_handle(qp, kkt="K2", **cfg):
    """Handle with the symbolic phase done (host only, no GPU needed)."""
    from madqp_jl_b200 import _lib
    from madqp_jl_b200.mpc import get_index_constraints
    ic = get_index_constraints(qp.lvar, qp.uvar, qp.lcon, qp.ucon)
    h = _lib.Handle(kkt_kind={"K2": _lib.KKT_K2, "NORMAL": _lib.KKT_NORMAL, "K2.5": _lib.KKT_K2_SCALED}[kkt], **cfg)
    if kkt != "NORMAL":
        h.set_pattern_k2(qp.nvar, qp.ncon, qp.H_rows + 1, qp.H_cols + 1, qp.A_rows + 1, qp.A_cols + 1, ic.ind_ineq + 1)
    else:
        h.set_pattern_normal(qp.nvar, qp.ncon, qp.A_rows + 1, qp.A_cols + 1, ic.ind_ineq + 1)
    h.set_bounds_index(ic.ind_lb + 1, ic.ind_ub + 1)
    h.analyze()
    return h, ic
