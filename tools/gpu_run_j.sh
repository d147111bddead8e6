#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
ALT=$PWD/gast-net-3dposeestimation_b200/csrc/alt/libgast_b200_nospread.so
timeout 300 python tools/tc_probe.py --perf > $O/r02_j_perf.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5 > $O/r02_j_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_j_bench_new.json 2> $O/r02_j_bench_new.err
GAST_B200_LIB=$ALT timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_j_bench_prev.json 2> $O/r02_j_bench_prev.err
timeout 300 python tools/train_step.py 20 > $O/r02_j_train_step.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1500 --csv --log-file $O/r02_j_train_launches.csv python tools/train_step.py 3 > $O/r02_j_train_ncu.log 2>&1
grep -v "epilogue per" $O/r02_j_perf.txt | cut -c1-400; tail -3 $O/r02_j_pytest.log; cut -c1-200 $O/r02_j_bench_new.json; cut -c1-200 $O/r02_j_bench_prev.json; cat $O/r02_j_train_step.txt | tail -2
