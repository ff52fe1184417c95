"""Structure statistics of the supernodal elimination tree of a workload (host only, development aid)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madqp_jl_b200.mpc import MPCSolver, FixedRegularization
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = tuple(float(x) for x in v.split(",")) if "," in v else int(float(v))
qps, kkt, name = bench.make_workload(wl, scale)
t = time.time()
s = MPCSolver(qps[0], kkt_system=kkt, regularization=FixedRegularization(1e-8, -1e-8), **kw)
s.prepare_host()
print("analysis", time.time() - t, "s")
h = s.h
first = h.symbolic("SUPER_FIRST"); par = h.symbolic("SUPER_PARENT"); rp = h.symbolic("SUPER_ROWPTR"); lev = h.symbolic("SUPER_LEVEL")
ns = len(par)
w = np.diff(first); hh = np.diff(rp)
print(s.info)
print("nsuper", ns, "w mean", w.mean(), "h mean", hh.mean(), "panel mean", (w * hh).mean())
def hist(x, edges):
    c, _ = np.histogram(x, edges)
    return dict(zip([f"<{e}" for e in edges[1:]], c.tolist()))
print("w hist", hist(w, [0, 2, 3, 5, 9, 17, 33, 65, 1000]))
print("h hist", hist(hh, [0, 9, 17, 33, 65, 129, 257, 513, 100000]))
print("panel hist", hist(w * hh, [0, 65, 129, 257, 513, 1025, 2049, 4097, 8193, 16385, 1 << 30]))
# height from leaves
height = np.zeros(ns, dtype=np.int64)
for i in range(ns):
    p = par[i]
    if p >= 0:
        height[p] = max(height[p], height[i] + 1)
print("max height", height.max())
hc = np.bincount(height)
print("supernodes per height", hc.tolist())
fl = (w.astype(float) * hh * hh)
fh = np.bincount(height, weights=fl)
print("h*h*w per height (M)", (fh / 1e6).round(1).tolist())
pn = np.bincount(height, weights=(w * hh).astype(float))
print("panel doubles per height (M)", (pn / 1e6).round(2).tolist())
print("max panel per height", [int((w * hh)[height == k].max()) for k in range(height.max() + 1)])
print("max w per height", [int(w[height == k].max()) for k in range(height.max() + 1)])
print("max h per height", [int(hh[height == k].max()) for k in range(height.max() + 1)])
nchild = np.bincount(par[par >= 0], minlength=ns)
print("children hist", np.bincount(nchild)[:12].tolist())
# number of distinct ancestor supernodes each supernode updates
rows = h.symbolic("SUPER_ROWS")
snof = np.repeat(np.arange(ns), w)
tot = 0; mx = 0
cnt = np.zeros(ns, dtype=np.int64)
for i in range(0, ns, max(1, ns // 20000)):
    r = rows[rp[i] + w[i]: rp[i + 1]]
    u = np.unique(snof[r])
    cnt[i] = len(u)
sel = cnt[::max(1, ns // 20000)]
print("targets per supernode (sampled): mean", sel.mean(), "max", sel.max())
print("levels(top) count", lev.max(), "top supernodes", int((lev > 0).sum()))
# ---- fan-in analysis (pull / left-looking view)
d_of = np.repeat(np.arange(ns), hh)                    # owner of each sn_rows entry
pos = np.arange(len(rows)) - np.repeat(rp[:-1], hh)     # position inside the panel
below = pos >= np.repeat(w, hh)
d_b = d_of[below]; t_b = snof[rows[below]]; pos_b = pos[below]
key = d_b.astype(np.int64) * ns + t_b
uk, idx, cntJ = np.unique(key, return_index=True, return_counts=True)
d_u = uk // ns; t_u = uk % ns
I_len = hh[d_u] - pos_b[idx]                            # rows from the first J row to the end
fma = I_len.astype(float) * cntJ * w[d_u]
print("update pairs", len(uk), "total pull FMA (M)", fma.sum() / 1e6)
nsrc = np.bincount(t_u, minlength=ns)
pull = np.bincount(t_u, weights=fma, minlength=ns)
print("sources per target: mean", nsrc.mean(), "max", nsrc.max(), "hist", hist(nsrc, [0, 1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 100000]))
print("max sources per height", [int(nsrc[height == k].max()) for k in range(height.max() + 1)])
print("max pull MFMA per height", [round(float(pull[height == k].max()) / 1e6, 3) for k in range(height.max() + 1)])
print("sum pull MFMA per height", [round(float(pull[height == k].sum()) / 1e6, 1) for k in range(height.max() + 1)])
own = (w.astype(float) * w * hh / 2)
print("sum own-factor MFMA per height", [round(float(own[height == k].sum()) / 1e6, 1) for k in range(height.max() + 1)])
# bytes read by pulls (source suffix rows x w_d doubles)
rd = I_len.astype(float) * w[d_u]
print("pull reads (M doubles)", rd.sum() / 1e6, "vs stored", (w * hh).sum() / 1e6)
nbv = (hh - w).astype(np.int64)
small = (w * hh <= 12288) & (nbv <= 256)
print("small", int(small.sum()), "big", int((~small).sum()), "C block doubles (packed, small only) M:", float((nbv * (nbv + 1) // 2)[small].sum()) / 1e6,
      "max nb small", int(nbv[small].max()), "max panel small", int((w * hh)[small].max()), "max h small", int(hh[small].max()))
warp = small & (w * hh <= 512) & (hh <= 96)
print("warp-mode", int(warp.sum()), "block-mode", int((small & ~warp).sum()))
