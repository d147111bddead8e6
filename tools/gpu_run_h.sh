#!/bin/bash
# round-2 GPU session H: corrections folded into the group accumulator + 4-stage A ring
cd "$(dirname "$0")/.."
O=gpurun_out
ALT=$PWD/gast-net-3dposeestimation_b200/csrc/alt/libgast_b200_nofold.so
timeout 300 python tools/tc_probe.py > $O/r02_h_probe_numerics.txt 2>&1
timeout 300 python tools/tc_probe.py --perf > $O/r02_h_perf.txt 2>&1
rm -f $O/mpjpe_delta.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_realtime.py -m gpu -q 2>&1 | tail -15 > $O/r02_h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_h_bench_fold.json 2> $O/r02_h_bench_fold.err
GAST_B200_LIB=$ALT timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_h_bench_nofold.json 2> $O/r02_h_bench_nofold.err
timeout 120 python tools/launch_times.py > $O/r02_h_launch_times.txt 2>&1
cat $O/r02_h_probe_numerics.txt | tail -24; grep -v "per chunk\|epilogue per" $O/r02_h_perf.txt; tail -4 $O/r02_h_pytest.log; cut -c1-200 $O/r02_h_bench_fold.json; cut -c1-200 $O/r02_h_bench_nofold.json; cat $O/mpjpe_delta.txt; tail -3 $O/r02_h_bench_fold.err
