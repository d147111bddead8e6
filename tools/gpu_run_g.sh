#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python tools/tc_probe.py --perf > $O/r02_g_perf.txt 2>&1
grep -v "per chunk\|epilogue per" $O/r02_g_perf.txt
