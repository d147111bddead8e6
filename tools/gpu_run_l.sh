#!/bin/bash
# round-2 GPU session L: validation of the committed build (single issuer, folded corrections, graphed trainer)
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/train_fixture_report.txt $O/train_grad_noise.txt $O/mpjpe_delta.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -40 > $O/r02_l_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r02_l_bench.json 2> $O/r02_l_bench.err
timeout 100 python tools/launch_times.py > $O/r02_l_launch_times.txt 2>&1
timeout 200 python tools/train_step.py 20 > $O/r02_l_train_step.txt 2>&1
timeout 200 python tools/train_step.py 20 graph >> $O/r02_l_train_step.txt 2>&1
timeout 300 python tools/stream_bench.py > $O/r02_l_stream_bench.txt 2>&1
tail -12 $O/r02_l_pytest.log; cut -c1-250 $O/r02_l_bench.json; tail -3 $O/r02_l_bench.err; cat $O/r02_l_launch_times.txt | tail -29; cat $O/r02_l_train_step.txt; cat $O/r02_l_stream_bench.txt | cut -c1-200
