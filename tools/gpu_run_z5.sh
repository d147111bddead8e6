#!/bin/bash
# round-2 GPU session Z5: the shipped binary with GAST_TC_F16=0 (tf32 + bf16 corrections everywhere): full GPU suite and the
# headline bench line next to the default (fp16 form) on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
GAST_TC_F16=0 timeout 300 python -m pytest tests -m gpu -q --timeout 120 > $O/r02_z5_pytest_tf32.log 2>&1; echo "pytest tf32 rc $?"; tail -2 $O/r02_z5_pytest_tf32.log
GAST_TC_F16=0 timeout 120 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_z5_bench_tf32.json 2> $O/r02_z5_bench_tf32.err; cut -c1-200 $O/r02_z5_bench_tf32.json
timeout 120 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_z5_bench_f16.json 2> $O/r02_z5_bench_f16.err; cut -c1-200 $O/r02_z5_bench_f16.json
