#!/bin/bash
# round-2 GPU session M: two MMA issuers with per-issuer operand barriers (strict timeouts: an earlier protocol hung)
cd "$(dirname "$0")/.."
O=gpurun_out
D=$PWD/gast-net-3dposeestimation_b200/csrc/alt/libgast_b200_dual.so
GAST_B200_LIB=$D timeout 60 python tools/tc_probe.py > $O/r02_m_numerics_dual.txt 2>&1; echo "numerics rc $?"
GAST_B200_LIB=$D timeout 90 python tools/tc_probe.py --perf > $O/r02_m_perf_dual.txt 2>&1; echo "perf rc $?"
GAST_B200_LIB=$D timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -8 > $O/r02_m_pytest_dual.log; echo "pytest rc $?"
GAST_B200_LIB=$D timeout 150 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_m_bench_dual.json 2> $O/r02_m_bench_dual.err; echo "bench rc $?"
timeout 150 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_m_bench_single.json 2> $O/r02_m_bench_single.err
GAST_B200_LIB=$D timeout 60 python tools/launch_times.py > $O/r02_m_launch_times_dual.txt 2>&1
tail -8 $O/r02_m_numerics_dual.txt; grep -v "epilogue per" $O/r02_m_perf_dual.txt | cut -c1-400; tail -4 $O/r02_m_pytest_dual.log; cut -c1-200 $O/r02_m_bench_dual.json; cut -c1-200 $O/r02_m_bench_single.json; tail -4 $O/r02_m_launch_times_dual.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 1000 --csv --log-file $O/r02_m_train_launches.csv python tools/train_step.py 2 > $O/r02_m_train_ncu.log 2>&1
timeout 100 python -m pytest tests/test_gpu_train.py -m gpu -q -k graphed --timeout 60 2>&1 | tail -3
