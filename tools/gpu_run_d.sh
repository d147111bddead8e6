#!/bin/bash
# round-2 GPU session D: pipelined A converters + MMA issue-order experiment, split-M wgrad, streaming throughput
cd "$(dirname "$0")/.."
O=gpurun_out
ALT=$PWD/gast-net-3dposeestimation_b200/csrc/alt/libgast_b200_nopipe.so
timeout 300 python tools/tc_probe.py > $O/r02_d_probe_numerics.txt 2>&1
timeout 300 python tools/tc_probe.py --perf > $O/r02_d_probe_perf.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -30 > $O/r02_d_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_d_bench_new.json 2> $O/r02_d_bench_new.err
GAST_B200_LIB=$ALT timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_d_bench_nopipe.json 2> $O/r02_d_bench_nopipe.err
timeout 120 python tools/launch_times.py > $O/r02_d_launch_times.txt 2>&1
timeout 300 python tools/stream_bench.py > $O/r02_d_stream_bench.txt 2>&1
tail -4 $O/r02_d_pytest.log; cat $O/r02_d_probe_perf.txt | grep -v "per chunk\|epilogue per"; cut -c1-300 $O/r02_d_bench_new.json; cut -c1-200 $O/r02_d_bench_nopipe.json; cat $O/r02_d_stream_bench.txt; tail -3 $O/r02_d_bench_new.err
