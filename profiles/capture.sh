#!/bin/bash
# ncu captures of the dominant kernel of every bench config (run on the GPU box through gpurun):
#   bash profiles/capture.sh            -> gpurun_out/p_*.ncu-rep, p_launches_*.csv, p_plain_*.log
# Each ncu command runs only after the same command exited 0 without ncu (B200_PROFILING.md).
set -u
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --ess-iters 0"
run() {  # config, kernel regex, launches to skip
  local cfg=$1 kern=$2 skip=$3
  $B --config $cfg > $OUT/p_plain_$cfg.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o $OUT/p_prof_$cfg \
      $B --config $cfg > $OUT/p_ncu_$cfg.log 2>&1
  echo "$cfg rc=$?"
}
run c2 k_dense_tc3 3
run c3 k_logistic_tc 12
run c4 k_nbody 3
run c5 k_small 4
for cfg in c2 c3 c5; do
  $B --config $cfg > $OUT/p_plain2_$cfg.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/p_launches_$cfg.csv \
      $B --config $cfg > /dev/null 2>&1
  echo "launches $cfg rc=$?"
done
