#!/bin/bash
# round-2 GPU session N: truncation split in the A converters, split-M kernel for the theta/phi vector gradients
cd "$(dirname "$0")/.."
O=gpurun_out
ALT=$PWD/gast-net-3dposeestimation_b200/csrc/alt/libgast_b200_rnsplit.so
rm -f $O/train_fixture_report.txt $O/train_grad_noise.txt $O/mpjpe_delta.txt
timeout 60 python tools/tc_probe.py > $O/r02_n_numerics.txt 2>&1
timeout 500 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > $O/r02_n_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r02_n_bench.json 2> $O/r02_n_bench.err
GAST_B200_LIB=$ALT timeout 150 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_n_bench_rnsplit.json 2> $O/r02_n_bench_rnsplit.err
timeout 100 python tools/train_step.py 20 > $O/r02_n_train_step.txt 2>&1
timeout 100 python tools/train_step.py 20 graph >> $O/r02_n_train_step.txt 2>&1
tail -12 $O/r02_n_numerics.txt; tail -8 $O/r02_n_pytest.log; cut -c1-220 $O/r02_n_bench.json; cut -c1-220 $O/r02_n_bench_rnsplit.json; cat $O/r02_n_train_step.txt; cat $O/mpjpe_delta.txt; tail -3 $O/r02_n_bench.err
