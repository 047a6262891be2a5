#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full of each kernel.
# usage (under gpurun): bash tools/gpu_capture.sh <tag>
tag=${1:-cap}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/launches.csv \
    python bench.py --seconds 2 --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_front|k_rate_loop|k_psy_front|k_psy_scan|k_bits" -s 24 -c 6 \
    -o $out/full -f python bench.py --seconds 2 --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_full.log 2>&1
tail -3 $out/pytest.log; cat $out/bench.json; tail -2 $out/bench.err
