#!/bin/bash
# Round-2 GPU job 16 (2 GPUs): deeper step-size pipeline (adaptLag 3, 4) of the fused run at the 8-GPU shard size; tests.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -k "fused or adapt or run_" > $OUT/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02p_pytest.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 profiles/ens_wait_probe_ranks.py 1000 20 > $OUT/r02p_wait_ranks2_L20.txt 2>&1; echo "rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 profiles/ens_wait_probe_ranks.py 1000 4 > $OUT/r02p_wait_ranks2_L4.txt 2>&1; echo "rc=$?"
for lag in 2 3 4; do EHMC_ADAPT_LAG=$lag timeout 200 python profiles/fused_probe.py 2000 20 2>&1 | cut -c1-62 > $OUT/r02p_fused_L20_lag$lag.txt; cat $OUT/r02p_fused_L20_lag$lag.txt; done
