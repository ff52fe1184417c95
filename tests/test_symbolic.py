"""CPU tests of the library's symbolic phase (host code, no GPU): bit-exact against the golden
fixtures and against the oracle's restatements of the same reference functions."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, analyze_handle
from madqp_jl_b200 import _lib
from madqp_jl_b200 import problems as P
from oracle import oracle as O


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "madipm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(madipm_[a-z0-9_]+)\s*\(", hdr)))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), set(declared) ^ set(_lib.EXPORTED_SYMBOLS)


def test_version_string():
    lib = _lib.load()
    assert b"sm_100a" in lib.madipm_version()


def test_bad_arguments_are_reported_not_thrown():
    h = _lib.Handle()
    with pytest.raises(_lib.MadIPMError) as e:
        h.analyze()                      # before set_pattern
    assert e.value.code == _lib.BAD_ARGUMENT
    with pytest.raises(_lib.MadIPMError):
        h.set_pattern_k2(2, 1, [], [], [5], [1], [])      # row index out of range
    h2 = _lib.Handle(kkt_kind=_lib.KKT_NORMAL)
    with pytest.raises(_lib.MadIPMError):
        h2.set_pattern_k2(2, 1, [], [], [1], [1], [])     # wrong kind for this handle


def test_numeric_calls_fail_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    h, _ = analyze_handle(P.simple_lp())
    with pytest.raises(_lib.MadIPMError) as e:
        h.vec_fill("X", 0.0)
    assert e.value.code == _lib.CUDA_ERROR and "no CPU fallback" in str(e.value)


def test_csr_map_and_normal_pattern_golden(golden):
    """S1/S2/S3 through the C ABI against the fixtures (0-based getters vs 1-based fixtures)."""
    for c in golden["sparse_cases"]:
        if "Cp" not in c or len(c["I"]) == 0:
            continue
        m, n = c["m"], c["n"]
        h = _lib.Handle(kkt_kind=_lib.KKT_NORMAL, ordering=_lib.ORDER_NATURAL)
        h.set_pattern_normal(n, m, c["I"], c["J"], [])
        assert (h.symbolic("A_ROWPTR") + 1).tolist() == c["Bp"]
        assert (h.symbolic("A_COLIND") + 1).tolist() == c["Bj"]
        assert (h.symbolic("A_CSR_MAP") + 1).tolist() == c["order"]
        assert (h.symbolic("KKT_COLPTR") + 1).tolist() == c["Cp"]
        assert (h.symbolic("KKT_ROWVAL") + 1).tolist() == c["Cj"]
        # product lists reproduce tril(A D A') exactly
        Bx = np.array(c["Bx"]); D = np.array(c["D"])
        pp, pa, pb, pk = (h.symbolic(k) for k in ("PROD_PTR", "PROD_A", "PROD_B", "PROD_K"))
        Cx = [float(sum((Bx[pa[t]] * D[pk[t]]) * Bx[pb[t]] for t in range(pp[e], pp[e + 1]))) for e in range(len(c["Cj"]))]
        assert Cx == c["Cx"]


@pytest.mark.parametrize("maker,kw", [(P.sparse_lp, dict(n=400, m=200)), (P.transportation_lp, dict(n_src=30, n_dst=40, n_arcs=300))])
def test_normal_pattern_matches_oracle_scan(maker, kw):
    """S3 by column lists == the reference's O(m nnz) scan (oracle restatement), bit for bit."""
    qp = maker(**kw)
    h, _ = analyze_handle(qp, "NORMAL")
    Ap, Aj, _ = O.coo_to_csr(qp.ncon, qp.nvar, (qp.A_rows + 1).astype(np.int32), (qp.A_cols + 1).astype(np.int32),
                             np.ones(qp.nnzj))
    Cp, Cj = O.build_normal_system(qp.ncon, qp.nvar, Ap, Aj)
    assert np.array_equal(h.symbolic("KKT_COLPTR") + 1, Cp) and np.array_equal(h.symbolic("KKT_ROWVAL") + 1, Cj)
    assert np.array_equal(h.symbolic("A_ROWPTR") + 1, Ap) and np.array_equal(h.symbolic("A_COLIND") + 1, Aj)


@pytest.mark.parametrize("qp", [P.inequality_qp(seed=0), P.maros_meszaros_like(n=120, m=60, nnz=400)])
def test_k2_pattern_matches_oracle(qp):
    """S4: aug COO block order + coo_to_csc map == the oracle's restatement."""
    h, ic = analyze_handle(qp, "K2")
    o = O.MPCOracle(qp)
    assert np.array_equal(h.symbolic("KKT_COLPTR") + 1, o.aug_colptr)
    assert np.array_equal(h.symbolic("KKT_ROWVAL") + 1, o.aug_rowval)
    assert np.array_equal(h.symbolic("AUG_CSC_MAP") + 1, o.aug_csc_map)
    assert np.array_equal(h.symbolic("A_CSR_MAP") + 1, o.A_csr_map)
    info = h.info()
    assert info["n_slack"] == ic.ind_ineq.size and info["kkt_order"] == qp.nvar + ic.ind_ineq.size + qp.ncon


@pytest.mark.parametrize("ordering", [_lib.ORDER_AMD, _lib.ORDER_ND, _lib.ORDER_NATURAL])
@pytest.mark.parametrize("maker,kw,kkt", [
    (P.banded_qp, dict(n=3000, m=1500), "K2"),
    (P.maros_meszaros_like, dict(n=300, m=150, nnz=1500), "K2"),
    (P.sparse_lp, dict(n=2000, m=1000), "NORMAL"),
    (P.simple_lp, {}, "K2"),
])
def test_symbolic_factorisation_is_exact(ordering, maker, kw, kkt):
    """Given the library's permutation, its etree and column counts equal the oracle's
    LDL symbolic phase bit for bit; the supernodal structure covers exactly the pattern of L."""
    qp = maker(**kw)
    h, _ = analyze_handle(qp, kkt, ordering=ordering, nd_leaf=64)
    info = h.info()
    N = info["kkt_order"]
    perm = h.symbolic("PERM")
    assert np.array_equal(np.sort(perm), np.arange(N))
    cp, ri = h.symbolic("KKT_COLPTR"), h.symbolic("KKT_ROWVAL")
    s = O.LDLSolver(N, (cp + 1).astype(np.int32), (ri + 1).astype(np.int32), np.ones(ri.size), perm)
    assert np.array_equal(h.symbolic("ETREE_PARENT"), s.parent)
    assert np.array_equal(h.symbolic("COLCOUNT"), s.colcount)
    assert info["nnz_L"] == s.nnz_L and info["factor_flops"] == s.flops
    # supernodes: partition of the columns, rows sorted, own columns first, exact cover of L's pattern
    first, rp, rows = h.symbolic("SUPER_FIRST"), h.symbolic("SUPER_ROWPTR"), h.symbolic("SUPER_ROWS")
    assert first[0] == 0 and first[-1] == N and np.all(np.diff(first) > 0)
    # pattern of L from the oracle's numeric factor of a diagonally dominant matrix with that pattern
    iperm = np.empty(N, dtype=np.int64); iperm[perm] = np.arange(N)
    col_of = np.repeat(np.arange(N), np.diff(cp))
    vals = np.where(ri == col_of, 4.0 * N, 1e-3)
    s2 = O.LDLSolver(N, (cp + 1).astype(np.int32), (ri + 1).astype(np.int32), vals, perm)
    assert s2.factorize()
    lcols = [set() for _ in range(N)]
    for j in range(N):
        lcols[j] = set(s2.Li[s2.Lp[j]:s2.Lp[j + 1]].tolist())
    for sn in range(first.size - 1):
        r = rows[rp[sn]:rp[sn + 1]]
        w = first[sn + 1] - first[sn]
        assert np.array_equal(r[:w], np.arange(first[sn], first[sn + 1])) and np.all(np.diff(r) > 0)
        for k, j in enumerate(range(first[sn], first[sn + 1])):
            assert lcols[j] <= set(r[k + 1:].tolist())      # relaxed supernodes may hold extra (zero) rows
    # assembly map: injective, and lands on the (row, col) it claims
    amap, xoff = h.symbolic("ASM_MAP"), h.symbolic("SUPER_XOFF")
    assert np.unique(amap).size == amap.size
    sn_of = np.repeat(np.arange(first.size - 1), np.diff(first))
    for p in np.random.default_rng(0).integers(0, amap.size, min(200, amap.size)):
        a, b = iperm[ri[p]], iperm[col_of[p]]
        r, c = max(a, b), min(a, b)
        sn = sn_of[c]
        hgt = rp[sn + 1] - rp[sn]
        off = amap[p] - xoff[sn]
        assert off // hgt == c - first[sn] and rows[rp[sn] + off % hgt] == r


def test_fill_is_comparable_to_scipy_minimum_degree():
    """Ordering quality sanity check: AMD / ND fill within 1.6x of SuperLU's MMD(A'+A)."""
    from scipy.sparse.linalg import splu
    qp = P.banded_qp(n=4000, m=2000)
    h, _ = analyze_handle(qp, "K2", ordering=_lib.ORDER_AMD)
    cp, ri = h.symbolic("KKT_COLPTR"), h.symbolic("KKT_ROWVAL")
    N = h.info()["kkt_order"]
    col_of = np.repeat(np.arange(N), np.diff(cp))
    K = sp.csc_matrix((np.where(ri == col_of, 1e3, 1.0), (ri, col_of)), shape=(N, N))
    K = K + sp.tril(K, -1).T
    lu = splu(K.tocsc(), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    assert h.info()["nnz_L"] <= 1.6 * lu.L.nnz
    h2, _ = analyze_handle(qp, "K2", ordering=_lib.ORDER_ND, nd_leaf=128)
    assert h2.info()["nnz_L"] <= 3.0 * lu.L.nnz
    assert h2.info()["n_levels"] < h.info()["n_levels"]          # the reason ND is the default


def test_schedule_covers_every_supernode_once():
    qp = P.banded_qp(n=6000, m=3000)
    h, _ = analyze_handle(qp, "K2", nd_leaf=100)
    level, parent, kind = h.symbolic("SUPER_LEVEL"), h.symbolic("SUPER_PARENT"), h.symbolic("SUPER_KIND")
    task, pptr, nblk = h.symbolic("TASK_SN"), h.symbolic("PHASE_PTR"), h.symbolic("PHASE_NBLOCK")
    info = h.info()
    assert info["n_levels"] == level.max() + 1 == len(pptr) - 1
    assert info["n_block_tasks"] + info["n_warp_tasks"] + info["n_big"] == info["nsuper"]
    for s, p in enumerate(parent):
        if p >= 0:
            assert p > s                                      # postorder
            assert level[p] > level[s]                        # a parent is processed in a later phase
    # every small supernode is exactly one task, in the phase of its height
    assert sorted(task.tolist()) == np.flatnonzero(kind != 2).tolist()
    cls = h.symbolic("PHASE_CLS").reshape(-1, 5)
    first, rp = h.symbolic("SUPER_FIRST"), h.symbolic("SUPER_ROWPTR")
    panel = np.diff(first) * np.diff(rp)
    for k in range(len(pptr) - 1):
        assert (level[task[pptr[k]:pptr[k + 1]]] == k).all()
        assert 0 <= nblk[k] <= pptr[k + 1] - pptr[k]
        assert cls[k, 0] == pptr[k] and cls[k, 4] == pptr[k + 1] and (np.diff(cls[k]) >= 0).all()
        assert cls[k, 3] - cls[k, 0] == nblk[k]
    used = [int((cls[:, c + 1] - cls[:, c]).sum()) for c in range(4)]
    assert sum(used) == len(task) and used[3] > 0 and sum(used[:3]) > 0      # warp tasks and multi-warp tasks both occur


@pytest.mark.parametrize("maker,kw,cfg", [
    (P.banded_qp, dict(n=6000, m=3000), dict(nd_leaf=100)),
    (P.banded_qp, dict(seed=3, n=20000, m=10000), dict()),
    (P.small_qp, dict(instance=0), dict()),
    (P.maros_meszaros_like, dict(seed=0, n=200, m=100, nnz=800), dict(ordering=_lib.ORDER_AMD)),
])
def test_ticket_order_is_a_dependency_order_for_the_dataflow_kernels(maker, kw, cfg):
    """k_tree_df (kernels_tree.cuh) hands out tickets in TASK_SN order (reversed by height for the backward sweep) and lets a
    ticket holder spin until its children (its parent) have signalled.  That cannot deadlock iff everything a task waits for
    has a smaller ticket: checked here on the host, together with the counts the kernel compares the counters against, by
    playing the schedule with a bounded number of groups and random task durations."""
    h, _ = analyze_handle(maker(**kw), "K2", **cfg)
    level, parent = h.symbolic("SUPER_LEVEL"), h.symbolic("SUPER_PARENT")
    cptr, cidx = h.symbolic("CHILD_PTR"), h.symbolic("CHILD_IDX")
    task, pptr = h.symbolic("TASK_SN"), h.symbolic("PHASE_PTR")
    nsuper = len(parent)
    task_of = np.full(nsuper, -1)
    task_of[task] = np.arange(len(task))
    small = task_of >= 0
    for s_ in range(nsuper):                                   # the child lists and the parent array say the same thing
        assert (parent[cidx[cptr[s_]:cptr[s_ + 1]]] == s_).all()
    assert np.array_equal(np.bincount(parent[parent >= 0], minlength=nsuper), np.diff(cptr))
    # forward (factor, forward sweep): children before parents
    kids_small = np.array([small[cidx[cptr[s_]:cptr[s_ + 1]]].all() for s_ in task])
    for q, s_ in enumerate(task):
        kids = cidx[cptr[s_]:cptr[s_ + 1]]
        if kids_small[q]:
            assert (task_of[kids] < q).all() and (level[kids] < level[s_]).all()
    # backward sweep: heights descending, so a parent's phase comes first
    order_bwd = np.concatenate([np.arange(pptr[k], pptr[k + 1]) for k in range(len(pptr) - 2, -1, -1)]) if len(task) else task
    pos_bwd = np.empty(len(task), dtype=np.int64)
    pos_bwd[order_bwd] = np.arange(len(task))
    for q, s_ in enumerate(task):
        if parent[s_] >= 0 and small[parent[s_]]:
            assert pos_bwd[task_of[parent[s_]]] < pos_bwd[q]
    if not kids_small.all():
        return                                                 # multi-block fronts: the dataflow launches are not used
    # play the forward schedule: G groups, tickets in order, a holder waits for its counter; must finish for any durations
    rng = np.random.default_rng(0)
    for G in (1, 3, 64):
        dur = rng.integers(1, 20, len(task))
        counter = np.zeros(len(task), dtype=np.int64)
        need = np.diff(cptr)[task]
        free_at = np.zeros(G, dtype=np.int64)                  # when each group takes its next ticket
        done_at = np.zeros(len(task), dtype=np.int64)
        for q, s_ in enumerate(task):                          # ticket q goes to the group that asks first
            g = int(np.argmin(free_at))
            kids = task_of[cidx[cptr[s_]:cptr[s_ + 1]]]
            ready = max([free_at[g]] + [done_at[k] for k in kids])         # every child already HAS a finishing time: no cycle
            done_at[q] = ready + dur[q]
            free_at[g] = done_at[q]
            if parent[s_] >= 0:
                counter[task_of[parent[s_]]] += 1
        assert np.array_equal(counter, need)                   # the value k_tree_df waits for (m.c1 - m.c0)


def _multifrontal_emulation(h, K):
    """numpy restatement of the data flow of kernels_tree.cuh (pull of packed contribution blocks through
    PAR_POS / SUPER_COFF) on the dense permuted matrix K; returns (L, d) with K = L diag(d) L'."""
    first, rp, rows = h.symbolic("SUPER_FIRST"), h.symbolic("SUPER_ROWPTR"), h.symbolic("SUPER_ROWS")
    cptr, cidx, ppos, coff = h.symbolic("CHILD_PTR"), h.symbolic("CHILD_IDX"), h.symbolic("PAR_POS"), h.symbolic("SUPER_COFF")
    task, pptr = h.symbolic("TASK_SN"), h.symbolic("PHASE_PTR")
    N = K.shape[0]
    Lf, dd = np.eye(N), np.zeros(N)
    C = np.zeros(coff[-1])

    def col(b, nb):
        return b * nb - b * (b + 1) // 2

    for k in range(len(pptr) - 1):
        for s in task[pptr[k]:pptr[k + 1]]:
            w = first[s + 1] - first[s]
            r = rows[rp[s]:rp[s + 1]]
            hh = len(r)
            nb = hh - w
            Pn = np.tril(K[np.ix_(r, r[:w])])                       # assembled panel (lower part of the diagonal block)
            Cs = np.zeros((nb, nb))
            for c in cidx[cptr[s]:cptr[s + 1]]:
                wc = first[c + 1] - first[c]
                nbc = rp[c + 1] - rp[c] - wc
                pos = ppos[rp[c] + wc: rp[c + 1]]
                assert (np.diff(pos) > 0).all() and pos[0] < w
                assert (rows[rp[c] + wc: rp[c + 1]] == r[pos]).all()
                for b in range(nbc):
                    for a in range(b, nbc):
                        v = C[coff[c] + col(b, nbc) + a]
                        if pos[b] < w:
                            Pn[pos[a], pos[b]] += v
                        else:
                            Cs[pos[a] - w, pos[b] - w] += v
            for j in range(w):                                        # dense LDL' of the panel
                d = Pn[j, j]
                u = Pn[j + 1:, j].copy()
                Pn[j + 1:, j] = u / d
                for kk in range(j + 1, w):
                    Pn[kk:, kk] -= u[kk - j - 1:] * Pn[kk, j]
            dd[r[:w]] = np.diag(Pn[:w, :w])
            for j in range(w):
                Lf[r[j + 1:], r[j]] = Pn[j + 1:, j]
            L21 = Pn[w:, :]
            Cs -= np.tril(L21 @ np.diag(dd[r[:w]]) @ L21.T)
            for b in range(nb):
                C[coff[s] + col(b, nb) + b: coff[s] + col(b, nb) + nb] = Cs[b:, b]
    return Lf, dd


@pytest.mark.parametrize("ordering", [_lib.ORDER_ND, _lib.ORDER_AMD])
def test_contribution_block_tables_reproduce_the_factor(ordering):
    """PAR_POS, SUPER_COFF, CHILD_*, TASK_SN and PHASE_PTR carry a complete multifrontal factorisation."""
    qp = P.banded_qp(n=120, m=60, half_width=8)
    h, _ = analyze_handle(qp, "K2", ordering=ordering, nd_leaf=24)
    assert (h.symbolic("SUPER_KIND") == 0).all()
    cp, ri, perm = h.symbolic("KKT_COLPTR"), h.symbolic("KKT_ROWVAL"), h.symbolic("PERM")
    N = h.info()["kkt_order"]
    rng = np.random.default_rng(0)
    col_of = np.repeat(np.arange(N), np.diff(cp))
    vals = np.where(ri == col_of, np.where(col_of < qp.nvar, 4.0, -4.0), rng.uniform(-1, 1, len(ri)))
    K = np.zeros((N, N))
    K[ri, col_of] = vals
    K = K + np.tril(K, -1).T
    Kp = K[np.ix_(perm, perm)]
    L, d = _multifrontal_emulation(h, Kp)
    assert np.abs(L @ np.diag(d) @ L.T - Kp).max() <= 1e-10 * np.abs(Kp).max()


def test_nested_dissection_does_not_depend_on_host_threads(monkeypatch):
    """csrc/dissection.cpp runs the halves of the first dissection levels on separate host threads; every region owns a
    fixed range of the order, so the permutation is the same for any thread count (and from run to run)."""
    qp = P.banded_qp(seed=3, n=60000, m=30000)
    perms = []
    for threads in ("1", "2", "16", None):
        if threads is None:
            monkeypatch.delenv("MADIPM_ND_THREADS", raising=False)
        else:
            monkeypatch.setenv("MADIPM_ND_THREADS", threads)
        h, _ = analyze_handle(qp, "K2", ordering=_lib.ORDER_ND)
        perms.append(h.symbolic("PERM"))
        h.close()
    assert sorted(perms[0].tolist()) == list(range(perms[0].size))
    for p in perms[1:]:
        assert np.array_equal(p, perms[0])


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/madipm_b200.h must compile as C99 (what a ccall / cgo / ctypes binding
    assumes), and a C program must link against the library and drive the host-side entry points."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "madipm_b200.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    src = tmp_path / "drive.c"
    src.write_text(r'''
#include <stdio.h>
#include "madipm_b200.h"
int main(void) {
    madipm_config cfg; madipm_handle* h = 0; madipm_info info; madipm_scale_options so;
    madipm_default_config(&cfg);
    cfg.kkt_kind = MADIPM_KKT_K2_SCALED;
    if (madipm_create(&cfg, &h) != MADIPM_OK) return 1;
    /* min x1 + x2  s.t. x1 + x2 = 1, x >= 0  (simple_lp of the reference's tests) */
    int32_t jI[2] = {1, 1}, jJ[2] = {1, 2};
    int64_t lb[2] = {1, 2};
    if (madipm_set_pattern_k2(h, 2, 1, 0, 0, 0, 2, jI, jJ, 0, 0) != MADIPM_OK) return 2;
    if (madipm_set_bounds_index(h, 2, lb, 0, 0) != MADIPM_OK) return 3;
    if (madipm_analyze(h) != MADIPM_OK) return 4;
    if (madipm_get_info(h, &info) != MADIPM_OK) return 5;
    madipm_default_scale_options(&so);
    printf("%s kind=%d order=%d nlb=%d sweeps=%d\n", madipm_version(), info.kkt_kind, info.kkt_order, info.nlb, so.max_iter);
    return madipm_destroy(h);
}
''')
    exe = tmp_path / "drive"
    lib_dir = os.path.join(root, "madqp.jl_b200")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-l:libmadipm_b200.so", "-Wl,-rpath," + lib_dir])
    out = subprocess.check_output([str(exe)], text=True)
    assert "kind=2 order=3 nlb=2 sweeps=10" in out, out


@pytest.mark.parametrize("maker,kw,kkt,cfg", [
    (P.transportation_lp, dict(seed=0, n_src=300, n_dst=300, n_arcs=12000), "NORMAL", dict(block_panel=2048)),
    (P.sparse_lp, dict(seed=1, n=6000, m=3000), "NORMAL", dict(block_panel=1024)),
    (P.banded_qp, dict(seed=0, n=8000, m=4000), "K2", dict(block_panel=512)),
])
def test_scatter_table_shortcuts_agree_with_the_plain_search(maker, kw, kkt, cfg, monkeypatch):
    """The scatter tables of the right-looking update (symbolic.cpp step 7b) take two shortcuts for nested row structures
    (whole-list coincidence; parent = child's below rows): MADIPM_CHECK_TABLES=1 re-derives every position by binary search
    inside analyze and throws on any difference.  Small block_panel forces multi-block supernodes, where the tables matter."""
    monkeypatch.setenv("MADIPM_CHECK_TABLES", "1")
    h, _ = analyze_handle(maker(**kw), kkt, **cfg)
    info = h.info()
    assert info["n_big"] > 0
    h.close()
