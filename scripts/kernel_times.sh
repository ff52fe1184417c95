#!/bin/bash
# device time of every launch of 20 IPM iterations (ncu launch list, cold-cache serialised: compare shares) -> per-kernel summary
# usage: scripts/kernel_times.sh <tag> [env assignments...]
tag=$1; shift
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-c5 > gpurun_out/launches_$tag.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_$tag.csv > gpurun_out/launches_${tag}_summary.txt
head -12 gpurun_out/launches_${tag}_summary.txt
