#!/bin/bash
# Round-2 GPU job 6 (final single-GPU evidence): full default bench line, reference arm, ncu captures of the fused
# ensemble run, fused probe.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py > $OUT/r02f_bench_n1.json 2> $OUT/r02f_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02f_bench_reference.json 2> $OUT/r02f_bench_reference.err; echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --ess-iters 0 --no-others --no-sustained"
for cfg in c5 c5l4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small_ens -s 1 -c 1 -f -o $OUT/p3_prof_$cfg $B --config $cfg > $OUT/p3_ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/p3_launches_c5.csv $B --config c5 > /dev/null 2>&1; echo "launches c5 rc=$?"
timeout 300 python profiles/fused_probe.py 2000 20 > $OUT/r02f_fused_probe_L20.txt 2>&1; echo "fused20 rc=$?"
timeout 300 python profiles/fused_probe.py 2000 4 > $OUT/r02f_fused_probe_L4.txt 2>&1; echo "fused4 rc=$?"
timeout 600 python profiles/soak.py > $OUT/r02f_soak.txt 2>&1; echo "soak rc=$?"
