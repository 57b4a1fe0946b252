#!/bin/bash
# compute-sanitizer over the kernels rebuilt in round 2 (GPU box): crowd kernel (double-buffered rows, per-warp
# queues, ego ring), specialised vehicle kernels, scenario windows, device union table
mkdir -p gpurun_out
out=gpurun_out/sanitizer_r02.txt
: > $out
run() {  # tool, pytest -k expression
  echo "== $1  -k \"$2\"" >> $out
  timeout 1200 compute-sanitizer --tool $1 --print-limit 3 python -m pytest tests -m gpu -x -q -k "$2" 2>&1 \
    | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|passed|failed|Error" | head -8 >> $out
}
run racecheck "cell_grid"
run racecheck "vehicles_vs_oracle or highway_rss or large_vehicle_groups"
run memcheck "cell_grid or big_group or scenario_windows or union_table or host_path_without"
run racecheck "scenario_windows and 3"
cat $out
