#!/bin/bash
# A/B of the fused forward's tunables on one wave of C4 sequences (development aid; bench.py is the contract)
mkdir -p gpurun_out
for cfg in "0 0" "4 0" "8 0" "0 4" "0 8" "8 8" "16 8" "64 0"; do
  set -- $cfg
  echo "PF_PRO=$1 PF_DIST=$2"
  SMCT_PF_PRO=$1 SMCT_PF_DIST=$2 python scripts/quick_time.py --B 148 --iters 3 --algo fused
done
