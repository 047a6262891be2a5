#!/bin/bash
# ncu --set full with source of ONE launch of one kernel: tools/ncu_one.sh <tag> <kernel regex> [bench args...]
tag=$1; rx=$2; shift 2
mkdir -p gpurun_out/$tag
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 4 -c 1 -o gpurun_out/$tag/full_$rx -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-variants --pipeline serial "$@" > gpurun_out/$tag/ncu_$rx.log 2>&1
tail -2 gpurun_out/$tag/ncu_$rx.log
