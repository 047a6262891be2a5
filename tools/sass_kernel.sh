#!/bin/bash
# Build the library and dump one kernel's SASS with source-line markers:  tools/sass_kernel.sh <kernel substring> [out]
# (CPU only: nvcc cross-compiles; use it to check loop bodies, register counts and spills before spending GPU time)
set -e
k=${1:-k_rate_loop}; out=${2:-/tmp/sass_$k.txt}
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; exit 1; }
tmp=$(mktemp -d); (cd $tmp && cuobjdump -xelf all "$OLDPWD/mp3-enc-bsd_b200/libmp3gpu.so" > /dev/null && nvdisasm --print-line-info mp3gpu.sm_100a.cubin > all.sass)
awk -v k="$k" '/\.section[ \t]/ {on = (index($0, ".text.") > 0 && index($0, k) > 0)} on' $tmp/all.sass > $out
cuobjdump -res-usage mp3-enc-bsd_b200/libmp3gpu.so 2>/dev/null | grep -A1 "$k" | grep -o "REG:[0-9]*\|STACK:[0-9]*\|SHARED:[0-9]*" | tr '\n' ' '; echo
echo "$(grep -cE '^[[:space:]]*/\*[0-9a-f]{4,}\*/' $out) SASS instructions in $out"
rm -rf $tmp
