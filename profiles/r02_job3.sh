#!/bin/bash
# Round-2 GPU job 3: parity after sub-batched queue items + API changes; queue-item size sweep; C5 legs.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02c_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $OUT/r02c_pytest.log
for ss in 0 1 2 3; do
  EHMC_ENS_SSHIFT=$ss timeout 300 python profiles/fused_probe.py 2000 20 > $OUT/r02c_fused_probe_L20_ss$ss.txt 2>&1; echo "fused20 ss=$ss rc=$?"
  EHMC_ENS_SSHIFT=$ss timeout 300 python profiles/fused_probe.py 2000 4 > $OUT/r02c_fused_probe_L4_ss$ss.txt 2>&1; echo "fused4 ss=$ss rc=$?"
done
for cfg in c5 c5l4; do
  timeout 300 python bench.py --config $cfg --no-cpu-baseline --no-e2e --ess-iters 0 --no-others --no-sustained > $OUT/r02c_bench_$cfg.json 2>&1; echo "bench $cfg rc=$?"
done
