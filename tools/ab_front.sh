#!/bin/bash
# A/B of front-end kernel builds / variants (run under gpurun): tools/ab_front.sh "<lib|-> <variant>" ...
# prints the front kernel's ms per step, its fraction of the measured HBM peak and the identical-frame fraction vs exact
for spec in "$@"; do
  set -- $spec
  lib=$1; v=$2
  if [ "$lib" = "-" ]; then unset MP3GPU_LIB; else export MP3GPU_LIB=$PWD/$lib; fi
  python bench.py --clips 4144 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-variants --front $v 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']['front_polyphase_mdct']; print('$lib $v front_ms=%.2f frac=%.4f value=%.0f'%(k['ms_per_step'],k['frac_hbm'],d['value']))"
done
