#!/bin/bash
# Round-2 GPU job 5: where the fused run waits at the 8-GPU shard size; getSamples probe.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python profiles/ens_wait_probe.py 1000 20 > $OUT/r02e_ens_wait_L20.txt 2>&1; echo "wait20 rc=$?"
timeout 300 python profiles/ens_wait_probe.py 1000 4 > $OUT/r02e_ens_wait_L4.txt 2>&1; echo "wait4 rc=$?"
timeout 300 python profiles/getsamples_probe.py > $OUT/r02e_getsamples_probe.txt 2>&1; echo "getsamples rc=$?"
timeout 300 python -m pytest tests -m gpu -q -k "fused or adapt or run_" > $OUT/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02e_pytest.log
