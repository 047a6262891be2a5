#!/bin/bash
# bench at several chunk sizes: tools/ab_chunks.sh <tag> <clips> <chunk> ...
tag=$1; clips=$2; shift 2
mkdir -p gpurun_out/$tag
for cf in "$@"; do
  timeout 300 python bench.py --clips $clips --chunk-frames $cf --steps 2 --warmup 3 --no-cpu-baseline --no-variants --no-parity $PIPE > gpurun_out/$tag/cf$cf.json 2> gpurun_out/$tag/cf$cf.err
  python -c "
import json; d=json.load(open('gpurun_out/$tag/cf$cf.json')); print('chunk $cf', round(d['value']), round(d['e2e']['value']), {k:round(x['ms_per_step'],1) for k,x in d['kernels'].items()}, d['e2e'].get('output_crc32_first8'))"
done
