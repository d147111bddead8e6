"""Device-side counterpart of the reference's tools/mpii_coco_h36m.py (import as device.tools.mpii_coco_h36m): 2D keypoint format conversion on the device
(csrc/pipeline.cuh kpt_convert_kernel; float32 arithmetic in numpy's evaluation order, bit-identical for
float32 input).  Same names and return values: (keypoints_h36m, valid_frames)."""
import numpy as np
import torch

from gast_b200 import pipeline as _P


def _convert(keypoints, mode):
    as_numpy = not isinstance(keypoints, torch.Tensor)
    k = torch.as_tensor(np.ascontiguousarray(keypoints, dtype=np.float32)).cuda() if as_numpy else keypoints
    out, valid = _P.keypoints_convert(k, mode)
    if as_numpy:
        return out.cpu().numpy(), np.where(valid.cpu().numpy() != 0)[0]
    return out, torch.nonzero(valid, as_tuple=False).flatten()


def coco_h36m(keypoints):
    return _convert(keypoints, _P.KPT_COCO_H36M)


def mpii_h36m(keypoints):
    return _convert(keypoints, _P.KPT_MPII_H36M)


def coco_h36m_toe_format(keypoints):
    assert len(keypoints.shape) == 3
    return _convert(keypoints, _P.KPT_COCO_H36M_TOE)
