#!/bin/bash
# Round-2 GPU job 4 (2 GPUs): the default bench line at N = 2 (all configs; config 5 through the fused run with the
# in-kernel all-reduce and 64/128-particle queue items), then the C5 leg alone for a longer look.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $OUT/r02d_bench_n2.json 2> $OUT/r02d_bench_n2.err; echo "bench n2 rc=$?"
tail -c 600 $OUT/r02d_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --no-cpu-baseline --no-e2e --ess-iters 0 --no-others --no-sustained > $OUT/r02d_bench_c5_n2.json 2>&1; echo "c5 n2 rc=$?"
