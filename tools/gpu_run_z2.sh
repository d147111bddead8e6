#!/bin/bash
# round-2 GPU session Z2: tiled attention mix, warp-autonomous streaming row dots, staged expand (kernels_hbm.cuh):
# parity suite with the new defaults, per-launch times of new vs old kernels (env switches), other shapes, bench line
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
GAST_MIX_MODE=2 GAST_ROWDOT_STREAM=1 timeout 400 python -m pytest tests -m gpu -q --timeout 120 > $O/r02_z2_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r02_z2_pytest.log
N="GAST_MIX_MODE=2 GAST_ROWDOT_STREAM=1"
for v in "new:$N" "oldmix:GAST_MIX_MODE=0 GAST_ROWDOT_STREAM=1" "oldrd:GAST_MIX_MODE=2 GAST_ROWDOT_STREAM=0" "mixst2:$N GAST_MIX_STAGES=2" "minb3:$N GAST_B200_LIB=$A/libgast_b200_minb3.so"; do
  n=${v%%:*}; e=${v#*:}
  env $e timeout 60 python tools/launch_times.py > $O/r02_z2_lt_$n.txt 2>&1
  echo "== $n: $(grep -E 'expand|global_mix|rowdot|sum' $O/r02_z2_lt_$n.txt | awk '{printf "%s %s | ", $2, $3}')"
done
for s in "cfg4:2048 17 64 3,3,3,3" "cfg5:4096 19 128 3,3,3"; do
  n=${s%%:*}; a=${s#*:}
  env $N timeout 90 python tools/launch_times.py $a > $O/r02_z2_lt_${n}_new.txt 2>&1
  GAST_MIX_MODE=0 GAST_ROWDOT_STREAM=0 GAST_EXPAND_STAGED=0 timeout 90 python tools/launch_times.py $a > $O/r02_z2_lt_${n}_old.txt 2>&1
  for k in new old; do echo "== $n $k: $(grep -E 'expand|global_mix|rowdot|sum' $O/r02_z2_lt_${n}_$k.txt | awk '{printf "%s %s | ", $2, $3}')"; done
done
env $N timeout 200 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_z2_bench.json 2> $O/r02_z2_bench.err; echo "bench rc $?"; cut -c1-220 $O/r02_z2_bench.json
env $N timeout 120 python tools/graph_probe.py > $O/r02_z2_graph_probe.txt 2>&1; tail -2 $O/r02_z2_graph_probe.txt
