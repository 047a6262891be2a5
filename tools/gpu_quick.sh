#!/bin/bash
# quick GPU pass: parity tests + one bench line (no CPU baseline).  usage (under gpurun): bash tools/gpu_quick.sh <tag> [extra bench args]
tag=${1:-q}; shift
mkdir -p gpurun_out/$tag
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/$tag/pytest.log 2>&1; tail -3 gpurun_out/$tag/pytest.log
timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/$tag/bench.json 2> gpurun_out/$tag/bench.err
python -c "
import json; d=json.load(open('gpurun_out/$tag/bench.json')); print(round(d['value']), round(d['e2e']['value']), {k:round(v['ms_per_step'],1) for k,v in d['kernels'].items()}, d['roofline']['frac'])"
