#!/bin/bash
# quick GPU check: parity tests + short device-only bench lines of the named workloads
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
for w in "$@"; do
  case $w in
    c3n) a="--workload c3 --no-rss";;
    *) a="--workload $w";;
  esac
  python bench.py $a --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err || tail -5 gpurun_out/q_$w.err
  python - "$w" <<'PY'
import json,sys
w=sys.argv[1]
try:
    j=json.loads(open(f"gpurun_out/q_{w}.json").read().strip().splitlines()[-1])
    print(w, "value %.4g"%j["value"], "frac %.3f"%j["roofline"]["frac"], "kernel_ms %.3f"%j["roofline"]["kernel_ms"], j["clocks"])
except Exception as e: print(w, "failed", e)
PY
done
