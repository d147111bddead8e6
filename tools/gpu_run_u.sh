#!/bin/bash
# round-2 GPU session U: CTA pair as the default -- full GPU test-suite, bench with other configs, training step; variants
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
B="--steps 20 --warmup 5 --no-other-configs --no-cpu-baseline"
rm -f $O/train_fixture_report.txt $O/train_grad_noise.txt $O/mpjpe_delta.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > $O/r02_u_pytest.log; echo "pytest rc $?"; tail -5 $O/r02_u_pytest.log
timeout 120 python bench.py $B > $O/r02_u_bench_main.json 2> $O/r02_u_bench_main.err
for v in trunc cg2shallow a3; do
  GAST_B200_LIB=$A/libgast_b200_$v.so timeout 120 python bench.py $B > $O/r02_u_bench_$v.json 2> $O/r02_u_bench_$v.err; echo "bench $v rc $?"
done
GAST_TC_CG=1 timeout 120 python bench.py $B > $O/r02_u_bench_cg1.json 2> $O/r02_u_bench_cg1.err
for f in main trunc cg2shallow a3 cg1; do echo "$f: $(grep -o '"value": [0-9.]*, .*"ms_per_step": [0-9.]*' $O/r02_u_bench_$f.json | cut -c1-160)"; done
timeout 100 python tools/train_step.py 20 graph > $O/r02_u_train_step.txt 2>&1
GAST_TC_CG=1 timeout 100 python tools/train_step.py 20 graph >> $O/r02_u_train_step.txt 2>&1
cat $O/r02_u_train_step.txt
GAST_TRAIN_TC=1 timeout 120 python tools/train_determinism.py 128 > $O/r02_u_determinism.txt 2>&1; grep -v Warning $O/r02_u_determinism.txt | grep "run-to-run\|after 3\|step losses"
GAST_B200_LIB=$A/libgast_b200_trunc.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -2
