#!/bin/bash
# round-2 GPU session V: CTA pair with a deeper raw-A ring (6 x 16 KB by TMA, B ring 4 x 16 KB) vs the previous split (4 / 6)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
B="--steps 20 --warmup 5 --no-other-configs --no-cpu-baseline"
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -2
timeout 120 python bench.py $B > $O/r02_v_bench_main.json 2> $O/r02_v_bench_main.err
for v in b3r4 b2r5 stearly; do
  GAST_B200_LIB=$A/libgast_b200_$v.so timeout 120 python bench.py $B > $O/r02_v_bench_$v.json 2> $O/r02_v_bench_$v.err; echo "bench $v rc $?"
done
for f in main b3r4 b2r5 stearly; do echo "$f: $(grep -o '"value": [0-9.]*, .*"ms_per_step": [0-9.]*' $O/r02_v_bench_$f.json | cut -c1-160)"; done
timeout 60 python tools/tc_probe.py --cg > $O/r02_v_cg.txt 2>&1; cat $O/r02_v_cg.txt
GAST_B200_LIB=$A/libgast_b200_stearly.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -2
