#!/bin/bash
# Round-2 GPU job 2: parity after the small-D kernel changes (stream version 2, packed HMC path), probes, C5 legs.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/r02b_pytest.log
timeout 300 python profiles/debug/mass_adapt_debug.py > $OUT/r02b_mass_debug.txt 2>&1; echo "mass rc=$?"
timeout 300 python profiles/fused_probe.py 2000 20 > $OUT/r02b_fused_probe_L20.txt 2>&1; echo "fused20 rc=$?"
timeout 300 python profiles/fused_probe.py 2000 4 > $OUT/r02b_fused_probe_L4.txt 2>&1; echo "fused4 rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --ess-iters 0 --no-others --no-sustained"
for cfg in c5 c5l4 c1; do
  timeout 300 python bench.py --config $cfg --no-cpu-baseline --no-e2e --ess-iters 0 --no-others --no-sustained > $OUT/r02b_bench_$cfg.json 2>&1; echo "bench $cfg rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small_ens -s 1 -c 1 -f -o $OUT/p2_prof_c5 $B --config c5 > $OUT/p2_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_small_ens -s 1 -c 1 -f -o $OUT/p2_prof_c5l4 $B --config c5l4 > $OUT/p2_ncu_c5l4.log 2>&1; echo "ncu c5l4 rc=$?"
timeout 300 python profiles/hbm_probe.py > $OUT/r02b_hbm_probe.txt 2>&1; echo "hbm rc=$?"
