#!/bin/bash
# round-2 GPU session R: CTA-pair MMAs (tcgen05 cta_group::2) -- numerics first, under a short timeout
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
B="--steps 20 --warmup 5 --no-other-configs --no-cpu-baseline"
GAST_TC_CG=2 timeout 60 python tools/tc_probe.py > $O/r02_r_numerics_cg2.txt 2>&1; rc=$?
echo "numerics cg2 rc $rc"; grep "tcgen05\|Error\|error" $O/r02_r_numerics_cg2.txt | head -24
if [ $rc -ne 0 ]; then tail -5 $O/r02_r_numerics_cg2.txt; nvidia-smi --query-gpu=name,memory.used --format=csv; exit 0; fi
GAST_TC_CG=2 timeout 60 python tools/tc_probe.py --train > $O/r02_r_train_gemms_cg2.txt 2>&1; echo "train gemms rc $?"; cat $O/r02_r_train_gemms_cg2.txt
GAST_TC_CG=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -6 > $O/r02_r_pytest_cg2.log; echo "pytest cg2 rc $?"; tail -4 $O/r02_r_pytest_cg2.log
GAST_TC_CG=2 timeout 120 python bench.py $B > $O/r02_r_bench_cg2.json 2> $O/r02_r_bench_cg2.err; echo "bench cg2 rc $?"
GAST_TC_CG=1 timeout 120 python bench.py $B > $O/r02_r_bench_cg1.json 2> $O/r02_r_bench_cg1.err
for f in cg1 cg2; do echo "$f: $(grep -o '"value": [0-9.]*, .*"ms_per_step": [0-9.]*' $O/r02_r_bench_$f.json | cut -c1-160)"; grep -o '"per_kernel_ms_event_pass": {[^}]*}' $O/r02_r_bench_$f.json; done
GAST_TC_CG=2 timeout 60 python tools/launch_times.py > $O/r02_r_launch_times_cg2.txt 2>&1; cat $O/r02_r_launch_times_cg2.txt | tail -30
