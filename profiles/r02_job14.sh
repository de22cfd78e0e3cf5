#!/bin/bash
# Round-2 GPU job 14 (N GPUs, N = $1): the default bench line.
set -u
N=$1
OUT=gpurun_out
mkdir -p $OUT
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > $OUT/r02n_bench_n$N.json 2> $OUT/r02n_bench_n$N.err; echo "bench n$N rc=$?"
tail -c 300 $OUT/r02n_bench_n$N.err
