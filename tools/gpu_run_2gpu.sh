#!/bin/bash
# round-2 two-GPU session (gpurun --gpus 2): bench line at N = 2 (weak scaling, other_configs incl. the training step with
# its NCCL all-reduce timed on NVLink) and the data-parallel training smoke (parameters identical across ranks)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_final_bench_2gpu.json 2> $O/r02_final_bench_2gpu.err; echo "bench2 rc $?"; cut -c1-200 $O/r02_final_bench_2gpu.json
timeout 200 $TR --master-port 29512 tools/ddp_smoke.py > $O/r02_final_ddp_smoke.txt 2>&1; echo "ddp rc $?"; tail -2 $O/r02_final_ddp_smoke.txt
