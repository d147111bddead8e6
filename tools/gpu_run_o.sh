#!/bin/bash
# round-2 GPU session O (re-entry): green baseline of head + training step launch list
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 500 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > $O/r02_o_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r02_o_bench.json 2> $O/r02_o_bench.err
timeout 100 python tools/train_step.py 20 > $O/r02_o_train_step.txt 2>&1
timeout 100 python tools/train_step.py 20 graph >> $O/r02_o_train_step.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_o_train_launches.csv python tools/train_step.py 1 > $O/r02_o_train_ncu.log 2>&1
tail -8 $O/r02_o_pytest.log; cut -c1-300 $O/r02_o_bench.json; cat $O/r02_o_train_step.txt; tail -3 $O/r02_o_bench.err; wc -l $O/r02_o_train_launches.csv
