#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 profiles/ens_wait_probe_ranks.py 1000 20 > $OUT/r02o_wait_ranks2_L20.txt 2>&1; echo "rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 profiles/ens_wait_probe_ranks.py 1000 4 > $OUT/r02o_wait_ranks2_L4.txt 2>&1; echo "rc=$?"
