#!/bin/bash
# round-2 GPU session P: training GEMMs on the tcgen05 core (3xTF32), TC-core variants (dual issue, truncation split),
# block-1 clip slabs, attention-mix variant
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
B="--steps 20 --warmup 5 --no-other-configs --no-cpu-baseline"
rm -f $O/train_fixture_report.txt $O/train_grad_noise.txt $O/mpjpe_delta.txt
timeout 90 python tools/tc_probe.py > $O/r02_p_numerics.txt 2>&1; echo "numerics rc $?"
timeout 90 python tools/tc_probe.py --train > $O/r02_p_train_gemms.txt 2>&1; echo "train gemms rc $?"
timeout 500 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > $O/r02_p_pytest.log; echo "pytest rc $?"
timeout 100 python tools/train_step.py 20 graph > $O/r02_p_train_step.txt 2>&1
GAST_TRAIN_TC=0 timeout 100 python tools/train_step.py 20 graph >> $O/r02_p_train_step.txt 2>&1
timeout 100 python tools/train_step.py 20 >> $O/r02_p_train_step.txt 2>&1
timeout 120 python bench.py $B > $O/r02_p_bench_main.json 2> $O/r02_p_bench_main.err
for v in dual dualtrunc trunc; do
  GAST_B200_LIB=$A/libgast_b200_$v.so timeout 120 python bench.py $B > $O/r02_p_bench_$v.json 2> $O/r02_p_bench_$v.err; echo "bench $v rc $?"
done
for s in 4 8 16; do
  GAST_BLOCK1_SLABS=$s timeout 120 python bench.py $B > $O/r02_p_bench_slab$s.json 2> $O/r02_p_bench_slab$s.err
done
GAST_BLOCK1_SLABS=8 GAST_BLOCK_ORDER=1 timeout 120 python bench.py $B > $O/r02_p_bench_slab8_gfirst.json 2> $O/r02_p_bench_slab8_gfirst.err
GAST_MIX_V=2 timeout 120 python bench.py $B > $O/r02_p_bench_mix2.json 2> $O/r02_p_bench_mix2.err
GAST_B200_LIB=$A/libgast_b200_dualtrunc.so timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -4 > $O/r02_p_pytest_dualtrunc.log; echo "pytest dualtrunc rc $?"
GAST_MIX_V=2 GAST_BLOCK1_SLABS=8 timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 2>&1 | tail -4 > $O/r02_p_pytest_mix2_slab8.log
grep "3xTF32\|FFMA" $O/r02_p_numerics.txt | head -12; cat $O/r02_p_train_gemms.txt; tail -8 $O/r02_p_pytest.log; cat $O/r02_p_train_step.txt
for f in main dual dualtrunc trunc slab4 slab8 slab16 slab8_gfirst mix2; do echo "$f: $(cut -c1-190 $O/r02_p_bench_$f.json | grep -o '"value": [0-9.]*, .*"ms_per_step": [0-9.]*')"; done
grep -o '"per_kernel_ms_event_pass": {[^}]*}' $O/r02_p_bench_main.json $O/r02_p_bench_slab8.json $O/r02_p_bench_mix2.json
tail -3 $O/r02_p_pytest_dualtrunc.log $O/r02_p_pytest_mix2_slab8.log; tail -2 $O/r02_p_bench_main.err
