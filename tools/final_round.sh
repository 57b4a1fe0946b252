mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -2 gpurun_out/t.log
R=r01e
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2>> gpurun_out/bench_$R.err
python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_${R}_c5.json 2>> gpurun_out/bench_$R.err
python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/bench_${R}_c2.json 2>> gpurun_out/bench_$R.err
python bench.py --no-rss --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${R}_c3_norss.json 2>> gpurun_out/bench_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg_vehicle_kernel -s 1 -c 1 -o gpurun_out/prof_${R}_c3 python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/prof_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sg_vehicle_kernel -s 1 -c 1 -o gpurun_out/prof_${R}_c5 python bench.py --workload c5 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline >> gpurun_out/prof_$R.log 2>&1
echo done
