#!/bin/bash
# time the named workloads with every variants/libsg_*.so in turn (GPU box; no result checks):
#   bash tools/try_variants.sh c3 [c5 ...]
cp scenario_gym_b200/csrc/libsg_b200.so /tmp/libsg_keep.so
for v in variants/libsg_*.so; do
  cp $v scenario_gym_b200/csrc/libsg_b200.so
  for w in "$@"; do echo "$v $(python tools/time_rollout.py $w 2>&1 | tail -1)"; done
done
cp /tmp/libsg_keep.so scenario_gym_b200/csrc/libsg_b200.so
