#!/bin/bash
# round-2 GPU session Z3: packed fp32x2 FMAs (fma.rn.f32x2) in the attention mix and the SemCH epilogue -- parity suite,
# per-launch times against a build without them on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
timeout 400 python -m pytest tests -m gpu -q --timeout 120 > $O/r02_z3_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r02_z3_pytest.log
for s in "cfg2:4096 17 128 3,3,3" "cfg4:2048 17 64 3,3,3,3" "cfg5:4096 19 128 3,3,3"; do
  n=${s%%:*}; a=${s#*:}
  timeout 90 python tools/launch_times.py $a > $O/r02_z3_lt_${n}_ffma2.txt 2>&1
  GAST_B200_LIB=$A/libgast_b200_noffma2.so timeout 90 python tools/launch_times.py $a > $O/r02_z3_lt_${n}_scalar.txt 2>&1
  for k in ffma2 scalar; do echo "== $n $k: $(grep -E 'semch|global_mix|sum' $O/r02_z3_lt_${n}_$k.txt | awk '{printf "%s %s | ", $2, $3}')"; done
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_z3_bench.json 2> $O/r02_z3_bench.err; echo "bench rc $?"; cut -c1-220 $O/r02_z3_bench.json
GAST_B200_LIB=$A/libgast_b200_noffma2.so timeout 200 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02_z3_bench_scalar.json 2> $O/r02_z3_bench_scalar.err; cut -c1-220 $O/r02_z3_bench_scalar.json
