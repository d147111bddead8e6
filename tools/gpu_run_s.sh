#!/bin/bash
# round-2 GPU session S: why the CTA-pair (cta_group::2) kernel is 2x slower: MMA-only floors, deep B ring
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
A=$PWD/gast-net-3dposeestimation_b200/csrc/alt
GAST_TC_CG=1 timeout 60 python tools/tc_probe.py --cg > $O/r02_s_cg.txt 2>&1; echo "cg1 rc $?"
GAST_TC_CG=2 timeout 60 python tools/tc_probe.py --cg >> $O/r02_s_cg.txt 2>&1; echo "cg2 deep rc $?"
GAST_TC_CG=2 GAST_B200_LIB=$A/libgast_b200_cg2shallow.so timeout 60 python tools/tc_probe.py --cg >> $O/r02_s_cg.txt 2>&1; echo "cg2 shallow rc $?"
cat $O/r02_s_cg.txt
GAST_TC_CG=2 timeout 40 python tools/tc_probe.py 2>&1 | grep "K=1536" | head -3
