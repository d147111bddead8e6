#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over a small end-to-end workload of the shipped library
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 60 python tools/sanitize_small.py > $O/r02_final_sanitize_plain.txt 2>&1; echo "plain rc $?"; tail -2 $O/r02_final_sanitize_plain.txt
timeout 200 compute-sanitizer --tool memcheck --launch-timeout 0 --print-limit 20 python tools/sanitize_small.py > $O/r02_final_sanitize_memcheck.txt 2>&1; echo "memcheck rc $?"; tail -4 $O/r02_final_sanitize_memcheck.txt
timeout 200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > $O/r02_final_sanitize_racecheck.txt 2>&1; echo "racecheck rc $?"; tail -4 $O/r02_final_sanitize_racecheck.txt
