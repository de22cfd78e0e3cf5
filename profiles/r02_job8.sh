#!/bin/bash
# Round-2 GPU job 8 (8 GPUs): the default bench line at N = 8 (C2 headline + C3/C4/C5/C5-L4 legs; config 5 through the
# fused run with the in-kernel all-reduce over 7 peer mailboxes).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > $OUT/r02g_bench_n8.json 2> $OUT/r02g_bench_n8.err; echo "bench n8 rc=$?"
tail -c 400 $OUT/r02g_bench_n8.err
