#!/bin/bash
# Run on the GPU box (gpurun): bench lines of every workload, the reference arm, the ncu launch list of
# the default bench command and one full ncu capture of the dominant kernel of each workload.
# usage: bash tools/profile_round.sh r01c
R=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 400 gpurun_out/bench_$R.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err
for w in c2 c4 c5; do
  python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_${R}_$w.json 2>> gpurun_out/bench_$R.err
done
python bench.py --no-rss --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_c3_norss.json 2>> gpurun_out/bench_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg_vehicle_kernel -s 1 -c 1 \
    -o gpurun_out/prof_${R}_c3 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/prof_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg_vehicle_kernel -s 1 -c 1 \
    -o gpurun_out/prof_${R}_c5 python bench.py --workload c5 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline >> gpurun_out/prof_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg_replay_kernel -s 1 -c 1 \
    -o gpurun_out/prof_${R}_c2 python bench.py --workload c2 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline >> gpurun_out/prof_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg_rollout_kernel -s 1 -c 1 \
    -o gpurun_out/prof_${R}_c4 python bench.py --workload c4 --scenarios-per-gpu 296 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline >> gpurun_out/prof_$R.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_$R.csv
echo done
