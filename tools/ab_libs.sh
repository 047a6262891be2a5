#!/bin/bash
# A/B of library builds: tools/ab_libs.sh <tag> <clips> lib1.so lib2.so ...   (paths relative to mp3-enc-bsd_b200/)
tag=$1; clips=$2; shift 2
mkdir -p gpurun_out/$tag
for lib in "$@"; do
  MP3GPU_LIB=$PWD/mp3-enc-bsd_b200/$lib timeout 300 python bench.py --clips $clips --steps 2 --warmup 3 --no-cpu-baseline --no-variants --no-parity --pipeline serial > gpurun_out/$tag/$lib.json 2> gpurun_out/$tag/$lib.err
  python -c "
import json; d=json.load(open('gpurun_out/$tag/$lib.json')); print('$lib', round(d['value']), round(d['e2e']['value']), {k:round(x['ms_per_step'],1) for k,x in d['kernels'].items()}, d['e2e'].get('output_crc32_first8'))"
done
