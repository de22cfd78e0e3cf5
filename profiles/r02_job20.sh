#!/bin/bash
# Round-2 GPU job 20 (one GPU): whole GPU suite once more, ncu capture of the N-body kernel (config 4), probes.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02q_pytest.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --ess-iters 0 --no-others --no-sustained"
timeout 300 $B --config c4 > $OUT/p5_plain_c4.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nbody -s 3 -c 1 -f -o $OUT/p5_prof_c4 $B --config c4 > $OUT/p5_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 300 python profiles/hbm_probe.py > $OUT/r02q_hbm_probe.txt 2>&1; echo "hbm rc=$?"
timeout 300 python profiles/fused_probe.py 2000 4 2>&1 | cut -c1-140 > $OUT/r02q_fused_probe_L4.txt; cat $OUT/r02q_fused_probe_L4.txt
timeout 300 python profiles/fused_probe.py 2000 20 2>&1 | cut -c1-140 > $OUT/r02q_fused_probe_L20.txt; cat $OUT/r02q_fused_probe_L20.txt
