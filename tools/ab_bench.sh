#!/bin/bash
# A/B runs of library variants: tools/ab_bench.sh "<lib> <streams>" ...   (run under gpurun)
for spec in "$@"; do
  set -- $spec
  MP3GPU_LIB=$PWD/$1 python bench.py --streams $2 --seconds 4 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$1 streams=$2 value=%.0f e2e=%.0f crc=%s | '%(d['value'],d['e2e']['value'],d['e2e'].get('output_crc32_first8'))+' '.join('%s=%.1f'%(n,v['ms_per_step']) for n,v in k.items()))"
done
