#!/bin/bash
# round-2 GPU session K: two MMA issuers (alternate flush groups) against one
cd "$(dirname "$0")/.."
O=gpurun_out
ALT=$PWD/gast-net-3dposeestimation_b200/csrc/alt/libgast_b200_single.so
timeout 300 python tools/tc_probe.py --perf > $O/r02_k_perf.txt 2>&1
timeout 300 python tools/tc_probe.py > $O/r02_k_numerics.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02_k_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_k_bench_dual.json 2> $O/r02_k_bench_dual.err
GAST_B200_LIB=$ALT timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_k_bench_single.json 2> $O/r02_k_bench_single.err
timeout 120 python tools/launch_times.py > $O/r02_k_launch_times.txt 2>&1
grep -v "epilogue per" $O/r02_k_perf.txt | cut -c1-400; tail -4 $O/r02_k_pytest.log; cut -c1-200 $O/r02_k_bench_dual.json; cut -c1-200 $O/r02_k_bench_single.json; tail -8 $O/r02_k_numerics.txt; tail -3 $O/r02_k_bench_dual.err
