#!/bin/bash
# compute-sanitizer over smoke() and two golden tests: memcheck, then racecheck (shared-memory hazards) and initcheck
tag=${1:-san}
mkdir -p gpurun_out/$tag
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/$tag/${tool}_smoke.log 2>&1
  tail -4 gpurun_out/$tag/${tool}_smoke.log
done
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/$tag/memcheck_tests.log 2>&1
tail -5 gpurun_out/$tag/memcheck_tests.log
