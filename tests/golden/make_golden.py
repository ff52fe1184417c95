"""Generates tests/golden/*.json.

The reference is Julia and cannot run here, and it ships no golden vectors, so
these fixtures are derived from the DEFINITIONS the reference code implements, by independent
dense / brute-force Python that shares no code with oracle/ or the product:

  coo_to_csr           (/root/reference/src/utils.jl:148-191)  = stable sort of the triplets by row
  build_normal_system  (utils.jl:199-264)                      = lower triangle of the boolean A A'
  assemble_normal_system! (utils.jl:266-298)                   = tril(A D A') on that pattern
  transfer!            (scripts/cuda_wrapper.jl:17-28)         = dest[map[k]] += src[k]
  get_alpha_max_*      (kernels.jl:242-288)                    = left fold with init (1.0, 0)
  simple_lp            (test/runtests.jl:24-55)                = objective 1.0 (analytic)
  _scale_augmented_system_coo_kernel! (scripts/cuda_wrapper.jl:90-111) = its branch list, entry by entry (K2.5 build)

Values are small integers / dyadic rationals so every product and sum is exact in binary64 and
the fixtures can be compared bit-for-bit.  Run:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def case_sparse(seed, m, n, nnz, dup=False):
    rng = np.random.default_rng(seed)
    lin = rng.choice(m * n, size=nnz, replace=False)
    rng.shuffle(lin)
    I, J = lin // n + 1, lin % n + 1          # 1-based, unsorted
    V = rng.integers(-4, 5, nnz).astype(float)
    V[V == 0] = 1.0
    if dup:   # duplicate (i, j) entries: coo_to_csr keeps both, in input order
        I = np.concatenate([I, I[:3]]); J = np.concatenate([J, J[:3]]); V = np.concatenate([V, [7.0, 8.0, 9.0]])
    return I, J, V


def golden_coo_to_csr(I, J, V, m):
    order = sorted(range(len(I)), key=lambda k: I[k])      # Python's sort is stable
    Bp = [1]
    for i in range(1, m + 1):
        Bp.append(Bp[-1] + int(np.sum(I == i)))
    return Bp, [int(J[k]) for k in order], [float(V[k]) for k in order], [k + 1 for k in order]


def golden_normal(I, J, V, m, n, D):
    A = np.zeros((m, n))
    for i, j, v in zip(I, J, V):
        A[i - 1, j - 1] = v
    pattern = (np.abs(A) > 0).astype(int)
    S = pattern @ pattern.T
    Cp, Cj = [1], []
    for i in range(m):
        rows = [j + 1 for j in range(i, m) if S[j, i] > 0]
        Cj += rows
        Cp.append(Cp[-1] + len(rows))
    C = A @ np.diag(D) @ A.T
    Cx = []
    for i in range(m):
        for c in range(Cp[i] - 1, Cp[i + 1] - 1):
            Cx.append(float(C[Cj[c] - 1, i]))
    return Cp, Cj, Cx


def fold_argmin(vals):
    acc = (1.0, 0)
    for i, v in enumerate(vals, start=1):
        e = (v, i)
        acc = acc if acc[0] < e[0] else e
    return acc


def golden_scale_aug_coo(I, J, V, scaling, n, m):
    """cuda_wrapper.jl:90-111, one entry at a time exactly as the kernel's branches read (1-based k, i, j)."""
    dest = [None] * len(V)
    for k in range(1, len(V) + 1):
        i, j = int(I[k - 1]), int(J[k - 1])
        if k <= n:                                             # primal regularisation pr_diag
            dest[k - 1] = float(V[k - 1])
        elif i <= n and j <= n:                                # Hessian block
            dest[k - 1] = float(V[k - 1] * scaling[i - 1] * scaling[j - 1])
        elif n + 1 <= i <= n + m and j <= n:                   # Jacobian block
            dest[k - 1] = float(V[k - 1] * scaling[j - 1])
        elif n + 1 <= i <= n + m and n + 1 <= j <= n + m:      # dual regularisation du_diag
            dest[k - 1] = float(V[k - 1])
    return dest


def main():
    out = {}
    cases = []
    for seed, (m, n, nnz, dup) in enumerate([(5, 7, 12, False), (8, 6, 20, False), (4, 9, 10, True), (6, 6, 1, False), (3, 4, 0, False)]):
        I, J, V = case_sparse(seed, m, n, nnz, dup)
        Bp, Bj, Bx, order = golden_coo_to_csr(I, J, V, m)
        c = dict(m=m, n=n, I=I.tolist(), J=J.tolist(), V=V.tolist(), Bp=Bp, Bj=Bj, Bx=Bx, order=order)
        if not dup:
            D = (2.0 ** np.random.default_rng(100 + seed).integers(-2, 3, n)).tolist()
            Cp, Cj, Cx = golden_normal(I, J, V, m, n, D)
            c.update(D=D, Cp=Cp, Cj=Cj, Cx=Cx)
        cases.append(c)
    out["sparse_cases"] = cases
    # transfer!
    rng = np.random.default_rng(7)
    src = rng.integers(-8, 9, 15).astype(float)
    tmap = rng.integers(1, 7, 15)
    dest = np.zeros(6)
    for k in range(15):
        dest[tmap[k] - 1] += src[k]
    out["transfer"] = dict(src=src.tolist(), map=tmap.tolist(), ndest=6, dest=dest.tolist())
    # ratio-test fold semantics (ties, nothing blocking, value exactly 1.0)
    folds = []
    for vals in ([float("inf")] * 4, [0.5, 0.25, 0.25, 3.0], [1.0, 2.0, 1.0], [2.0, 3.0], [], [0.75], [1.0]):
        v, i = fold_argmin(vals)
        folds.append(dict(vals=[("inf" if np.isinf(x) else x) for x in vals], value=v, index=i))
    out["argmin_fold"] = folds
    # K2.5: aug COO in the block order pr_diag | tril(H) | J (rows + n) | slack | du_diag (cuda_wrapper.jl:95-107)
    n, m = 4, 3
    rng = np.random.default_rng(11)
    hI, hJ = [1, 2, 3, 3, 4, 4], [1, 1, 2, 3, 1, 4]                       # lower triangle incl. some diagonal entries
    jI, jJ = [1, 1, 2, 2, 3, 3], [1, 3, 2, 4, 1, 4]
    I = list(range(1, n + 1)) + hI + [i + n for i in jI] + [n + 2] + list(range(n + 1, n + m + 1))
    J = list(range(1, n + 1)) + hJ + jJ + [4] + list(range(n + 1, n + m + 1))   # one slack-like extra Jacobian entry (row 2, col 4)
    V = rng.integers(-6, 7, len(I)).astype(float)
    V[V == 0] = 2.0
    scaling = (2.0 ** rng.integers(-3, 4, n)).tolist()                    # dyadic: every product is exact
    out["scale_aug_coo"] = dict(n=n, m=m, I=I, J=J, V=V.tolist(), scaling=scaling,
                                dest=golden_scale_aug_coo(I, J, V, scaling, n, m),
                                source="/root/reference/scripts/cuda_wrapper.jl:90-111")
    out["simple_lp"] = dict(objective=1.0, solution=[0.5, 0.5], source="/root/reference/test/runtests.jl:24-55")
    with open(os.path.join(HERE, "reference_definitions.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote reference_definitions.json")


if __name__ == "__main__":
    main()
