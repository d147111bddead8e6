#!/bin/bash
# round-2 GPU session B: whole GPU suite on the bf16-correction kernel
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/train_fixture_report.txt $O/train_grad_noise.txt $O/mpjpe_delta.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/r02_b_pytest.log
tail -30 $O/r02_b_pytest.log; cat $O/train_fixture_report.txt
