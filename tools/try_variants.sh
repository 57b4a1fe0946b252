#!/bin/bash
# bench the named workload with every variants/libsg_*.so in turn (GPU box):  bash tools/try_variants.sh c3 [c5 ...]
mkdir -p gpurun_out
cp scenario_gym_b200/csrc/libsg_b200.so /tmp/libsg_keep.so
for v in variants/libsg_*.so; do
  cp $v scenario_gym_b200/csrc/libsg_b200.so
  for w in "$@"; do
    python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$w', 'value %.4g' % j['value'], 'kernel_ms %.3f' % j['roofline']['kernel_ms'])"
  done
done
cp /tmp/libsg_keep.so scenario_gym_b200/csrc/libsg_b200.so
