#!/bin/bash
# Freeze a copy of the tree under .stage/ so that a queued gpurun call runs a consistent snapshot while the
# working tree keeps changing (gpurun snapshots /root/repo when a box is acquired, not when the call is made).
# Usage: tools/stage.sh ; gpurun -- 'cd .stage && bash tools/<script>.sh'   (outputs go to ../gpurun_out)
set -e
cd "$(dirname "$0")/.."
rm -rf .stage
mkdir -p .stage gpurun_out
cp -r gast-net-3dposeestimation_b200 tests tools oracle include profiles bench.py __graft_entry__.py .stage/
[ -d baseline ] && cp -r baseline .stage/
[ -f MEASURED_PEAKS.json ] && cp MEASURED_PEAKS.json .stage/
find .stage -name __pycache__ -prune -exec rm -rf {} +
ln -s ../gpurun_out .stage/gpurun_out
echo "staged $(du -sh .stage | cut -f1)"
