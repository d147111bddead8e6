#!/bin/bash
# round-2 GPU session T: CTA pair with CTA-scope remote arrives; attribution of the slow variant
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
B="--steps 20 --warmup 5 --no-other-configs --no-cpu-baseline"
GAST_TC_CG=2 timeout 60 python tools/tc_probe.py --cg > $O/r02_t_cg.txt 2>&1; echo "cg2 rc $?"
GAST_TC_CG=2 GAST_B200_LIB=$A/libgast_b200_relcluster.so timeout 60 python tools/tc_probe.py --cg >> $O/r02_t_cg.txt 2>&1; echo "cg2 release.cluster rc $?"
cat $O/r02_t_cg.txt
GAST_TC_CG=2 timeout 40 python tools/tc_probe.py 2>&1 | grep "K=1536" | head -3
GAST_TC_CG=2 timeout 120 python bench.py $B > $O/r02_t_bench_cg2.json 2> $O/r02_t_bench_cg2.err; echo "bench cg2 rc $?"
GAST_TC_CG=1 timeout 120 python bench.py $B > $O/r02_t_bench_cg1.json 2> $O/r02_t_bench_cg1.err
for f in cg1 cg2; do echo "$f: $(grep -o '"value": [0-9.]*, .*"ms_per_step": [0-9.]*' $O/r02_t_bench_$f.json | cut -c1-160)"; grep -o '"per_kernel_ms_event_pass": {[^}]*}' $O/r02_t_bench_$f.json; done
GAST_TC_CG=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -3
GAST_TC_CG=2 timeout 60 python tools/launch_times.py > $O/r02_t_launch_times_cg2.txt 2>&1; tail -29 $O/r02_t_launch_times_cg2.txt
