#!/bin/bash
# development aid: time every experimental library build on one wave of C4 sequences
for f in smc-t-v2_b200/_exp/lib_*.so; do
  echo "== $f"
  SMCT_LIB_OVERRIDE=$PWD/$f timeout 120 python scripts/quick_time.py --B 148 --iters 3 --algo fused 2>&1 | tail -1 | cut -c1-120
done
