#!/bin/bash
# Round-2 GPU job 1: parity tests, the default bench line, ncu captures of the dominant kernels (one GPU).
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/r02_box.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/r02_pytest.log
timeout 600 python bench.py > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference.json 2> $OUT/r02_bench_reference.err; echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --ess-iters 0 --no-others --no-sustained"
run() {  # config, kernel regex, launches to skip
  local cfg=$1 kern=$2 skip=$3
  timeout 300 $B --config $cfg > $OUT/p_plain_$cfg.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o $OUT/p_prof_$cfg \
      $B --config $cfg > $OUT/p_ncu_$cfg.log 2>&1
  echo "$cfg rc=$?"
}
run c2 k_dense_tc3 3
run c3 k_logistic_tcs 12
run c5 k_small_ens 1
run c5l4 k_small_ens 1
for cfg in c2 c3 c5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/p_launches_$cfg.csv \
      $B --config $cfg > /dev/null 2>&1
  echo "launches $cfg rc=$?"
done
timeout 300 python profiles/fused_probe.py 2000 20 > $OUT/r02_fused_probe_L20.txt 2>&1; echo "fused20 rc=$?"
timeout 300 python profiles/fused_probe.py 2000 4 > $OUT/r02_fused_probe_L4.txt 2>&1; echo "fused4 rc=$?"
timeout 300 python profiles/hbm_probe.py > $OUT/r02_hbm_probe.txt 2>&1; echo "hbm rc=$?"
