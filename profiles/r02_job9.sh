#!/bin/bash
# Round-2 GPU job 9 (one GPU): the two-rank fused-run tests after the parallel mailbox receive, and an ncu capture of the
# fused run at the 8-GPU shard size (2^19 particles).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -q -k "fused or adapt or run_" > $OUT/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02h_pytest.log
timeout 200 python profiles/ens_ncu_target.py 19 20 300 > $OUT/r02h_plain_19.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small_ens -s 1 -c 1 -f -o $OUT/p4_prof_c5_2p19 python profiles/ens_ncu_target.py 19 20 300 > $OUT/p4_ncu_19.log 2>&1; echo "ncu 2^19 rc=$?"
