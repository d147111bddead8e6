#!/bin/bash
# round-2 GPU session E: ncu evidence for the current core -- launch list of the bench command, --set full of one
# forward of every BASELINE inference shape (27f/17j/128ch, 81f/17j/64ch, 27f/19j/128ch) with source-level stalls.
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --clock-control none"
# (1) launch list of the bench command (shares of the step)
timeout 600 $NCU --metrics gpu__time_duration.sum -s 200 -c 60 --csv --log-file $O/r02_e_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline > $O/r02_e_bench_under_ncu.log 2>&1
# (2) --set full of the 27 launches of one 4096-clip forward (warm-up: 3 forwards = 81 launches)
timeout 900 $NCU --set full --import-source on -k regex:"gemm_tc_kernel|global_mix|expand_kernel|rowdot8|shrink" -s 81 -c 27 \
    -o $O/r02_e_full_cfg2 python tools/launch_times.py 4096 17 128 3,3,3 > $O/r02_e_ncu_cfg2.log 2>&1
# (3) the other shapes: metrics only
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max"
timeout 600 $NCU --metrics $M -k regex:"gemm_tc_kernel|global_mix|expand_kernel|rowdot8|shrink" -s 108 -c 36 --csv \
    --log-file $O/r02_e_raw_cfg4.csv python tools/launch_times.py 2048 17 64 3,3,3,3 > $O/r02_e_ncu_cfg4.log 2>&1
timeout 600 $NCU --metrics $M -k regex:"gemm_tc_kernel|global_mix|expand_kernel|rowdot8|shrink" -s 81 -c 27 --csv \
    --log-file $O/r02_e_raw_cfg5.csv python tools/launch_times.py 8192 19 128 3,3,3 > $O/r02_e_ncu_cfg5.log 2>&1
ls -la $O | grep r02_e
