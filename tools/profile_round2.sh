#!/bin/bash
# Run on the GPU box (gpurun): the default bench line (all workloads as sub-records), the reference arm,
# the ncu launch list of the default command's headline part, and one full ncu capture of the dominant
# kernel of every workload AT ITS BENCHMARKED SIZE (so dram bytes / fp64 instructions are per launch of
# the launch the bench times).   usage: bash tools/profile_round2.sh r02
R=${1:-r02}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
tail -c 300 gpurun_out/bench_$R.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-subs > gpurun_out/launches_$R.log 2>&1
for w in c3 c5 c2 c4; do
  case $w in c3|c5) k=sg_vehicle_kernel;; c2) k=sg_replay_kernel;; c4) k=sg_crowd_kernel;; esac
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_${R}_$w \
      python bench.py --workload $w --steps 1 --warmup 2 --no-e2e --no-cpu-baseline --no-subs > gpurun_out/prof_${R}_$w.log 2>&1
  tail -1 gpurun_out/prof_${R}_$w.log
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_$R.csv
echo done
