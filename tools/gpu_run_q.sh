#!/bin/bash
# round-2 GPU session Q: run-to-run determinism of the training step, tcgen05 (3xTF32) vs FFMA GEMMs
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
GAST_TRAIN_TC=1 timeout 120 python tools/train_determinism.py 32 > $O/r02_q_determinism.txt 2>&1
GAST_TRAIN_TC=0 timeout 120 python tools/train_determinism.py 32 >> $O/r02_q_determinism.txt 2>&1
GAST_TRAIN_TC=1 timeout 120 python tools/train_determinism.py 128 >> $O/r02_q_determinism.txt 2>&1
cat $O/r02_q_determinism.txt
