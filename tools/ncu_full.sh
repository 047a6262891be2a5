#!/bin/bash
# ncu --set full of selected kernels at a wave-aligned batch: tools/ncu_full.sh <tag> <kernel regex> [streams]
tag=$1; rx=$2; streams=${3:-3552}
mkdir -p gpurun_out/$tag
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 12 -c 3 -o gpurun_out/$tag/full -f \
  python bench.py --streams $streams --seconds 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/$tag/ncu.log 2>&1
tail -3 gpurun_out/$tag/ncu.log
