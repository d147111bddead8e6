#!/bin/bash
# round-2 GPU session A: new tests on the round-1 kernel (known good), then numerics + tests + bench of the bf16-correction kernel
cd "$(dirname "$0")/.."
O=gpurun_out
ALT=$PWD/gast-net-3dposeestimation_b200/csrc/alt/libgast_b200_3xtf32.so
GAST_B200_LIB=$ALT timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r02_a_pytest_oldkernel.log
timeout 300 python tools/tc_probe.py > $O/r02_a_probe_numerics.txt 2>&1
timeout 300 python tools/tc_probe.py --perf > $O/r02_a_probe_perf.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_realtime.py -m gpu -q 2>&1 | tail -40 > $O/r02_a_pytest_newkernel.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_a_bench_new.json 2> $O/r02_a_bench_new.err
GAST_B200_LIB=$ALT timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_a_bench_old.json 2> $O/r02_a_bench_old.err
timeout 120 python tools/launch_times.py > $O/r02_a_launch_times_new.txt 2>&1
tail -5 $O/r02_a_pytest_oldkernel.log; tail -5 $O/r02_a_pytest_newkernel.log; tail -12 $O/r02_a_probe_numerics.txt; cut -c1-600 $O/r02_a_bench_new.json; tail -3 $O/r02_a_bench_new.err
