#!/bin/bash
# A/B of the psy_front FFT implementations + parity tests: bash tools/gpu_ab_psy.sh <tag> [clips]
tag=${1:-ab}; clips=${2:-4144}
mkdir -p gpurun_out/$tag
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$tag/pytest.log 2>&1; tail -3 gpurun_out/$tag/pytest.log
for v in regs program; do
  MP3GPU_PSY_FFT=$v timeout 300 python bench.py --clips $clips --steps 2 --warmup 3 --no-cpu-baseline --no-variants --parity-clips 8 --pipeline serial > gpurun_out/$tag/bench_$v.json 2> gpurun_out/$tag/bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/$tag/bench_$v.json')); print('$v', round(d['value']), round(d['e2e']['value']), {k:round(x['ms_per_step'],1) for k,x in d['kernels'].items()}, d.get('parity',{}).get('identical_frame_fraction'), d['e2e'].get('output_crc32_first8'))"
done
