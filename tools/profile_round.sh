#!/bin/bash
# Run on the GPU box (gpurun): bench line, ncu launch list and one full capture of the top kernel.
# usage: bash tools/profile_round.sh r01
R=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 600 gpurun_out/bench_$R.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg_vehicle_kernel -s 1 -c 1 \
    -o gpurun_out/prof_$R python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/prof_$R.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_$R.csv
echo done
