#!/bin/bash
# Round-2 capture on one GPU: parity tests, the default bench line (configs[3] as written), ncu launch list of the same
# command, ncu --set full (with source) of one launch of each hot kernel.  usage (under gpurun): bash tools/gpu_capture2.sh <tag>
tag=${1:-cap}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" >> $out/pytest.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(psy|front|rate|roll|bits|deint|call|seg|quant)" -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-variants --pipeline serial > $out/ncu_launch.log 2>&1
for k in k_psy_front_regs k_psy_scan k_front_tile k_rate_loop k_bits_emit; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 4 -c 1 -o $out/full_$k -f \
    python bench.py --clips 4144 --seconds 2 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-variants --pipeline serial > $out/ncu_$k.log 2>&1
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --parity-clips 1000 > $out/bench_parity1000.json 2> $out/bench_parity1000.err
tail -3 $out/pytest.log; cat $out/bench.json | head -c 1500; tail -2 $out/bench.err
