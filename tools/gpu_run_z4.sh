#!/bin/bash
# round-2 GPU session Z4: the main product on fp16 operands (PREC = 2, GAST_TC_F16=1): per-GEMM error against fp64,
# GEMM times, parity suite, per-launch times and bench line with the switch on, against the default on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 120 python tools/tc_probe.py --f16 > $O/r02_z4_f16_probe.txt 2>&1; echo "probe rc $?"; tail -16 $O/r02_z4_f16_probe.txt
GAST_TC_F16=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_loss.py tests/test_gpu_zz_realtime.py -m gpu -q --timeout 120 > $O/r02_z4_pytest_f16.log 2>&1; echo "pytest f16 rc $?"; tail -4 $O/r02_z4_pytest_f16.log
for v in "f16:GAST_TC_F16=1" "tf32:GAST_TC_F16=0"; do
  n=${v%%:*}; e=${v#*:}
  env $e timeout 60 python tools/launch_times.py > $O/r02_z4_lt_$n.txt 2>&1
  echo "== $n: $(grep -E 'gemm_tc|sum' $O/r02_z4_lt_$n.txt | awk '{printf "%s ", $3}')"
done
GAST_TC_F16=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_z4_bench_f16.json 2> $O/r02_z4_bench_f16.err; echo "bench rc $?"; cut -c1-200 $O/r02_z4_bench_f16.json
