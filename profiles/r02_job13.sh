#!/bin/bash
# Round-2 GPU job 13 (one GPU): groups-per-iteration sweep of the fused ensemble run.
set -u
OUT=gpurun_out
mkdir -p $OUT
for g in 256 512 1024 2048 4096; do
  EHMC_ENS_GROUPS=$g timeout 200 python profiles/fused_probe.py 2000 20 > $OUT/r02m_fused_L20_g$g.txt 2>&1; echo "g=$g rc=$?"; cut -c1-60 $OUT/r02m_fused_L20_g$g.txt
  EHMC_ENS_GROUPS=$g timeout 200 python profiles/fused_probe.py 2000 4 > $OUT/r02m_fused_L4_g$g.txt 2>&1; cut -c1-60 $OUT/r02m_fused_L4_g$g.txt
done
