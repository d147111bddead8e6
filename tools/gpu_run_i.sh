#!/bin/bash
# round-2 GPU session I: commits/probes spread between the MMAs of a chunk
cd "$(dirname "$0")/.."
O=gpurun_out
ALT=$PWD/gast-net-3dposeestimation_b200/csrc/alt/libgast_b200_nospread.so
timeout 300 python tools/tc_probe.py --perf > $O/r02_i_perf.txt 2>&1
GAST_B200_LIB=$ALT timeout 300 python tools/tc_probe.py --perf > $O/r02_i_perf_nospread.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_realtime.py -m gpu -q 2>&1 | tail -15 > $O/r02_i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_i_bench_spread.json 2> $O/r02_i_bench_spread.err
GAST_B200_LIB=$ALT timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_i_bench_nospread.json 2> $O/r02_i_bench_nospread.err
timeout 120 python tools/launch_times.py > $O/r02_i_launch_times.txt 2>&1
grep -v "epilogue per" $O/r02_i_perf.txt | cut -c1-400; echo; grep -v "per chunk\|epilogue per" $O/r02_i_perf_nospread.txt; tail -4 $O/r02_i_pytest.log; cut -c1-200 $O/r02_i_bench_spread.json; cut -c1-200 $O/r02_i_bench_nospread.json; tail -3 $O/r02_i_bench_spread.err
