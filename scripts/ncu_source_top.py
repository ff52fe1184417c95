"""Top source lines (stall samples / instructions) of one kernel from `ncu --page source --csv` (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
sections, cur, hdr = [], None, None
lastfile = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        lastfile = r[1].split('/')[-1]
    if len(r) >= 2 and r[0] == "Function Name":
        cur = {'name': r[1] + ' @ ' + (lastfile or ''), 'rows': []}; sections.append(cur)
    elif len(r) > 10 and r[0] == "Line No":
        hdr = r
    elif cur is not None and len(r) > 10:
        cur['rows'].append(r)
iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0
for s in sections:
    if pat not in s['name']:
        continue
    src = [r for r in s['rows'] if r[0] != '']
    tot_s = sum(num(r[iS]) for r in src); tot_i = sum(num(r[iI]) for r in src)
    print(s['name'][:80], "total samples", tot_s, "inst", tot_i)
    agg = {}
    for r in src:
        k = (int(r[0]), r[1])
        a = agg.setdefault(k, [0, 0]); a[0] += num(r[iS]); a[1] += num(r[iI])
    for (ln, text), (sm, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
        print(str(ln).rjust(4), "%5.1f%%" % (100 * sm / max(tot_s, 1)), "%5.1f%%" % (100 * ins / max(tot_i, 1)), text[:120])
