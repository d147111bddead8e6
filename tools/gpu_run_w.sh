#!/bin/bash
# round-2 GPU session W: evidence for the final core (CTA pair) -- full GPU test-suite, bench line, ncu launch list of the
# bench command, ncu --set full of one forward of configs[1], metrics of the 81f/64ch and 19j shapes, training step
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
rm -f $O/train_fixture_report.txt $O/train_grad_noise.txt $O/mpjpe_delta.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > $O/r02_w_pytest.log; echo "pytest rc $?"; tail -3 $O/r02_w_pytest.log
timeout 500 python bench.py --steps 20 --warmup 5 > $O/r02_w_bench.json 2> $O/r02_w_bench.err; echo "bench rc $?"; cut -c1-260 $O/r02_w_bench.json
timeout 60 python tools/launch_times.py > $O/r02_w_launch_times.txt 2>&1
timeout 100 python tools/train_step.py 20 graph > $O/r02_w_train_step.txt 2>&1; timeout 100 python tools/train_step.py 20 >> $O/r02_w_train_step.txt 2>&1; cat $O/r02_w_train_step.txt
timeout 120 python tools/stream_bench.py > $O/r02_w_stream_bench.json 2> $O/r02_w_stream_bench.err; echo "stream rc $?"
# (1) launch list of the bench command (shares of the step)
timeout 300 $NCU --metrics gpu__time_duration.sum -s 200 -c 60 --csv --log-file $O/r02_w_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline > $O/r02_w_bench_under_ncu.log 2>&1; echo "ncu launches rc $?"
# (2) --set full of the 27 launches of one 4096-clip forward (warm-up: 3 forwards = 81 launches)
timeout 600 $NCU --set full --import-source on -k regex:"gemm_tc_kernel|global_mix|expand_kernel|rowdot8|shrink" -s 81 -c 27 \
    -o $O/r02_w_full_cfg2 python tools/launch_times.py 4096 17 128 3,3,3 > $O/r02_w_ncu_cfg2.log 2>&1; echo "ncu full rc $?"
# (3) the other shapes: metrics only
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max"
timeout 400 $NCU --metrics $M -k regex:"gemm_tc_kernel|global_mix|expand_kernel|rowdot8|shrink" -s 108 -c 36 --csv \
    --log-file $O/r02_w_raw_cfg4.csv python tools/launch_times.py 2048 17 64 3,3,3,3 > $O/r02_w_ncu_cfg4.log 2>&1; echo "ncu cfg4 rc $?"
timeout 400 $NCU --metrics $M -k regex:"gemm_tc_kernel|global_mix|expand_kernel|rowdot8|shrink" -s 81 -c 27 --csv \
    --log-file $O/r02_w_raw_cfg5.csv python tools/launch_times.py 8192 19 128 3,3,3 > $O/r02_w_ncu_cfg5.log 2>&1; echo "ncu cfg5 rc $?"
# (4) training step launch list
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/r02_w_train_launches.csv python tools/train_step.py 1 > $O/r02_w_train_ncu.log 2>&1; echo "ncu train rc $?"
ls -la $O | grep r02_w
