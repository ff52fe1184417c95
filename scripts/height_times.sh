#!/bin/bash
# per-height device times of one factorisation (+ fused forward) and of the sweeps: every height its own launch (MADIPM_TREE_SPLIT=1), ncu launch list
# usage: scripts/height_times.sh <tag> [env assignments...]
tag=$1; shift
env MADIPM_TREE_SPLIT=1 "$@" ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 400 --csv --log-file gpurun_out/heights_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-c5 > gpurun_out/heights_$tag.log 2>&1
python - <<PY
import csv, collections
lines=[l for l in open('gpurun_out/heights_$tag.csv') if l.startswith('"')]
seq=[(r['Kernel Name'].split('(')[0].replace('void ','').replace('madipm::',''), float(r['Metric Value'].replace(',',''))/1000, r['Grid Size']) for r in csv.DictReader(lines)]
i=0; shown=collections.Counter()
while i<len(seq):
    j=i
    while j<len(seq) and seq[j][0]==seq[i][0]: j+=1
    nm=seq[i][0]
    if ('k_tree' in nm or 'k_front' in nm) and shown[nm]<1 and j-i>3:
        shown[nm]+=1
        print('$tag', nm, j-i, 'launches, total %.1f us:'%sum(s[1] for s in seq[i:j]), [round(s[1],1) for s in seq[i:j]])
    i=j
PY
