#!/bin/bash
# Round-2 GPU job 10 (8 GPUs): config 5 and its L = 4 variant through the fused run after the parallel mailbox receive.
set -u
OUT=gpurun_out
mkdir -p $OUT
for cfg in c5 c5l4; do
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --config $cfg --no-cpu-baseline --no-e2e --ess-iters 0 --no-others --no-sustained --steps 400 > $OUT/r02i_bench_${cfg}_n8.json 2> $OUT/r02i_bench_${cfg}_n8.err; echo "bench $cfg n8 rc=$?"
done
