#!/bin/bash
# round-2 GPU session C: whole GPU suite (bf16-correction kernel, O(1) streaming, unchanged callers)
cd "$(dirname "$0")/.."
O=gpurun_out
rm -f $O/train_fixture_report.txt $O/train_grad_noise.txt $O/mpjpe_delta.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -80 > $O/r02_c_pytest.log
tail -40 $O/r02_c_pytest.log; grep -E "^==|^step|^param" $O/train_fixture_report.txt
