#!/bin/bash
# Copies the reference's CALLER files (never its model/ -- that is what this repository replaces) into
# baseline/_ref/ so that tests/test_unchanged_callers.py can run reconstruction.reconstruction() and main.train()
# UNCHANGED over the drop-in on the GPU box, where /root/reference does not exist.  baseline/_ref/ is git-ignored
# (no reference source enters the history) but travels with gpurun.  Run in the build container.
set -e
REF=${1:-/root/reference}
cd "$(dirname "$0")/.."
D=baseline/_ref
rm -rf $D && mkdir -p $D/common $D/tools $D/data/keypoints $D/data/video
cp $REF/reconstruction.py $REF/main.py $D/
cp $REF/common/*.py $D/common/
cp $REF/tools/utils.py $REF/tools/mpii_coco_h36m.py $REF/tools/visualization.py $D/tools/
cp $REF/data/keypoints/baseball.json $D/data/keypoints/
cp $REF/data/video/baseball.mp4 $D/data/video/
# the drop-in owns these two modules of the `common` namespace package (same contract, tests compare them)
rm -f $D/common/graph_utils.py $D/common/skeleton.py
chmod -R u+w $D
echo "reference callers -> $D ($(du -sh $D | cut -f1))"
