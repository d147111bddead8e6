#!/bin/bash
# round-2 final GPU session: full GPU suite, bench line (all configs), launch list under ncu, ncu --set full of one forward
# (the report stays on the box: only its raw-metric and source-page exports come back, gpurun_out is limited to 64 MiB)
cd "$(dirname "$0")/.."
O=gpurun_out
T=${1:-r02_final}
mkdir -p $O
NCU="ncu --clock-control none"
timeout 400 python -m pytest tests -m gpu -q --timeout 120 > $O/${T}_gpu_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/${T}_gpu_pytest.log
timeout 500 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc $?"; cut -c1-200 $O/${T}_bench.json
timeout 60 python tools/launch_times.py > $O/${T}_launch_times.txt 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -s 200 -c 60 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline > $O/${T}_bench_under_ncu.log 2>&1; echo "ncu launches rc $?"
timeout 600 $NCU --set full --import-source on -k regex:"gemm_tc_kernel|global_mix|expand|rowdot8|shrink" -s 81 -c 27 \
    -o /tmp/${T}_full_cfg2 python tools/launch_times.py 4096 17 128 3,3,3 > $O/${T}_ncu_cfg2.log 2>&1; echo "ncu full rc $?"
ncu -i /tmp/${T}_full_cfg2.ncu-rep --page raw --csv > $O/${T}_raw_cfg2.csv 2>/dev/null
for id in 1 2 26; do
  ncu -i /tmp/${T}_full_cfg2.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $id --launch-count 1 2>/dev/null | gzip > $O/${T}_src_$id.csv.gz
done
ls -la $O | grep ${T}; du -sh $O
