#!/bin/bash
# development aid: build experimental variants of the library with -D switches, e.g.
#   scripts/dev/build_variants.sh base "" nopf "-DSMCT_X_PREFETCH=0"
cd "$(dirname "$0")/../.."
mkdir -p smc-t-v2_b200/_exp
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
    -o smc-t-v2_b200/_exp/lib_$name.so smc-t-v2_b200/csrc/smct_capi.cu -Xptxas -v 2>&1 | grep -A2 "fused_tc_forward" | grep -E "spill|registers" | tr '\n' ' ' &
done
wait
echo
ls -la smc-t-v2_b200/_exp/
