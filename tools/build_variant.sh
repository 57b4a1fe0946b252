#!/bin/bash
# build variants/libsg_<name>.so with one translation unit recompiled under extra flags:
#   bash tools/build_variant.sh <name> <unit.cu> [-DFLAG ...]
set -e
name=$1; unit=$2; shift 2
C=scenario_gym_b200/csrc
mkdir -p variants /tmp/variants_obj
obj=/tmp/variants_obj/${name}_${unit%.cu}.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xcompiler -fPIC -std=c++17 "$@" -c -o $obj $C/$unit
objs=""
for o in $C/build/*.o; do
  if [ "$(basename $o)" == "${unit%.cu}.o" ]; then objs="$objs $obj"; else objs="$objs $o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libsg_$name.so $objs
echo built variants/libsg_$name.so
