"""SASS / ptxas evidence for the shipped library (no GPU needed): per kernel, the counts of the mnemonics that identify the
hardware paths (DMMA = FP64 tensor core, LDGSTS = cp.async, UBLKCP / UTMALDG = TMA, HMMA / UTC*MMA = other tensor paths) and
the register / spill / shared-memory figures of `-Xptxas -v` (madqp.jl_b200/csrc/build.log).  Output: profiles/sass_rN.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "madqp.jl_b200", "libmadipm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = {"DMMA": r"\bDMMA\b", "LDGSTS": r"\bLDGSTS\b", "UBLKCP": r"\bUBLKCP\b", "UTMALDG": r"\bUTMALDG\b", "HMMA": r"\bHMMA\b",
       "UTCxMMA": r"\bUTC\w*MMA\b", "BAR": r"\bBAR\b", "ATOM(G).F64": r"\b(?:ATOMG?|RED)\.E\.ADD\.F64"}
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("madipm::", "")
        per[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        per[cur]["instructions"] += 1
        for k, p in pat.items():
            if re.search(p, line):
                per[cur][k] += 1
regs = {}
log = open(os.path.join(ROOT, "madqp.jl_b200", "csrc", "build.log")).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(?:, used (\d+) barriers)?", log):
    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("madipm::", "")
    regs[name] = (int(m.group(5)), int(m.group(3)), int(m.group(4)), int(m.group(6) or 0))
cols = list(pat)
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} | per-kernel mnemonic counts; ptxas -v figures from csrc/build.log")
print(f"# totals: " + ", ".join(f"{k} {sum(c[k] for c in per.values())}" for k in cols))
print(f"{'kernel':46s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in cols) + f" {'regs':>5s} {'spill_st':>8s} {'spill_ld':>8s} {'barriers':>8s}")
for name, c in sorted(per.items(), key=lambda kv: -kv[1]["instructions"]):
    r = regs.get(name, ("", "", "", ""))
    print(f"{name[:46]:46s} {c['instructions']:7d} " + " ".join(f"{c[k]:8d}" for k in cols) + f" {r[0]:>5} {r[1]:>8} {r[2]:>8} {r[3]:>8}")
