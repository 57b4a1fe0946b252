#!/bin/bash
# one full ncu capture of the dominant kernel of a workload at its benchmarked size
#   usage: bash tools/profile_one.sh c4 tag
w=$1; R=${2:-x}
mkdir -p gpurun_out
case $w in c3|c5) k=sg_vehicle_kernel;; c2) k=sg_replay_kernel;; c4) k=sg_crowd_kernel;; esac
ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_${R}_$w \
    python bench.py --workload $w --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-subs > gpurun_out/prof_${R}_$w.log 2>&1
tail -1 gpurun_out/prof_${R}_$w.log
