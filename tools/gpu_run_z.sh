#!/bin/bash
# round-2 GPU session Z: streaming expand / attention-mix kernels (kernels_hbm.cuh) -- parity suite, per-launch times of
# the new and the old kernels (env switches), ring-depth variants of the mix, other shapes
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q -x --timeout 120 > $O/r02_z_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r02_z_pytest.log
for v in "new:" "old:GAST_MIX_STREAM=0 GAST_EXPAND_STAGED=0" "mix1:GAST_MIX_STREAM=1" "mix3:GAST_MIX_STREAM=3"; do
  n=${v%%:*}; e=${v#*:}
  env $e timeout 60 python tools/launch_times.py > $O/r02_z_lt_$n.txt 2>&1
  echo "== $n: $(grep -E 'expand|global_mix|sum' $O/r02_z_lt_$n.txt | tr '\n' ' ')"
done
for s in "cfg4:2048 17 64 3,3,3,3" "cfg5:4096 19 128 3,3,3"; do
  n=${s%%:*}; a=${s#*:}
  timeout 90 python tools/launch_times.py $a > $O/r02_z_lt_${n}_new.txt 2>&1
  GAST_MIX_STREAM=0 GAST_EXPAND_STAGED=0 timeout 90 python tools/launch_times.py $a > $O/r02_z_lt_${n}_old.txt 2>&1
  for k in new old; do echo "== $n $k: $(grep -E 'expand|global_mix|sum' $O/r02_z_lt_${n}_$k.txt | tr '\n' ' ')"; done
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_z_bench.json 2> $O/r02_z_bench.err; echo "bench rc $?"; cut -c1-220 $O/r02_z_bench.json
