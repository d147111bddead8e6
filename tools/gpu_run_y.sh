#!/bin/bash
# round-2 GPU session Y: final binary -- bench line, launch list under ncu, ncu --set full of one forward (the report stays on the
# box: only its raw-metric and source-page exports come back, gpurun_out is limited to 64 MiB)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
timeout 500 python bench.py --steps 20 --warmup 5 > $O/r02_y_bench.json 2> $O/r02_y_bench.err; echo "bench rc $?"; cut -c1-200 $O/r02_y_bench.json
timeout 60 python tools/launch_times.py > $O/r02_y_launch_times.txt 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -s 200 -c 60 --csv --log-file $O/r02_y_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline > $O/r02_y_bench_under_ncu.log 2>&1; echo "ncu launches rc $?"
timeout 600 $NCU --set full --import-source on -k regex:"gemm_tc_kernel|global_mix|expand_kernel|rowdot8|shrink" -s 81 -c 27 \
    -o /tmp/r02_y_full_cfg2 python tools/launch_times.py 4096 17 128 3,3,3 > $O/r02_y_ncu_cfg2.log 2>&1; echo "ncu full rc $?"
ncu -i /tmp/r02_y_full_cfg2.ncu-rep --page raw --csv > $O/r02_y_raw_cfg2.csv 2>/dev/null
for id in 1 25; do
  ncu -i /tmp/r02_y_full_cfg2.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $id --launch-count 1 2>/dev/null | gzip > $O/r02_y_src_$id.csv.gz
done
ls -la $O | grep r02_y; du -sh $O
