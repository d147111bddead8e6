#!/bin/bash
# round-2 GPU session X: SemCH epilogue with a k-major coefficient slab and the self term from registers
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
B="--steps 20 --warmup 5 --no-other-configs --no-cpu-baseline"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -3
timeout 120 python bench.py $B > $O/r02_x_bench_main.json 2> $O/r02_x_bench_main.err
GAST_B200_LIB=$A/libgast_b200_oldsemch.so timeout 120 python bench.py $B > $O/r02_x_bench_oldsemch.json 2> $O/r02_x_bench_oldsemch.err
for f in main oldsemch; do echo "$f: $(grep -o '"value": [0-9.]*, .*"ms_per_step": [0-9.]*' $O/r02_x_bench_$f.json | cut -c1-160)"; grep -o '"per_kernel_ms_event_pass": {[^}]*}' $O/r02_x_bench_$f.json; done
timeout 60 python tools/launch_times.py > $O/r02_x_launch_times.txt 2>&1; grep semch $O/r02_x_launch_times.txt
GAST_B200_LIB=$A/libgast_b200_oldsemch.so timeout 60 python tools/launch_times.py 2>&1 | grep semch
