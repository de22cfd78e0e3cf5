#!/bin/bash
# Round-2 GPU job 11 (one GPU): smoke(), the full-size config 4 test, the whole GPU suite, the default bench line.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python __graft_entry__.py --smoke > $OUT/r02k_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/r02k_smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r02k_pytest.log
timeout 900 python bench.py > $OUT/r02k_bench_n1.json 2> $OUT/r02k_bench_n1.err; echo "bench rc=$?"
