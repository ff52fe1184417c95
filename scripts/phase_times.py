"""Per-height durations of the k_tree launches from an ncu launch list taken with MADIPM_TREE_SPLIT=1 (development aid)."""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
out = {}
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1000 if u in ('ns', 'nsecond') else v * 1000 if u in ('ms', 'msecond') else v
    op = re.search(r'k_tree<\D*(\d)>', row['Kernel Name'])
    out.setdefault(op.group(1) if op else row['Kernel Name'], []).append((round(v, 1), int(row['Grid Size'].strip('()').split(',')[0])))
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 29
for op, l in sorted(out.items()):
    l = l[:nh * (len(l) // nh)]
    print("op", op, "launches", len(l))
    for k in range(0, len(l), nh):
        seg = l[k:k + nh]
        print("  us:", [x[0] for x in seg], "sum", round(sum(x[0] for x in seg), 1))
        print("  grid:", [x[1] for x in seg])
