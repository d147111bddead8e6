#!/bin/bash
# round-2 GPU session F: what bounds the chunk rate -- B-ring depth / cluster coupling experiments
cd "$(dirname "$0")/.."
O=gpurun_out
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
for v in main c1 b4r2 c1b4r2; do
  if [ $v = main ]; then unset GAST_B200_LIB; else export GAST_B200_LIB=$A/libgast_b200_$v.so; fi
  timeout 300 python tools/tc_probe.py --perf > $O/r02_f_perf_$v.txt 2>&1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_f_bench_$v.json 2> $O/r02_f_bench_$v.err
  echo "== $v"; grep -v "per chunk\|epilogue per" $O/r02_f_perf_$v.txt; cut -c1-160 $O/r02_f_bench_$v.json
done
unset GAST_B200_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
