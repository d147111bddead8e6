"""Device-side counterpart of the reference's common/loss.py (same names and signatures; lives under
`device.` so that it does not shadow the reference module, which has more functions): `mpjpe` (loss.py:5-11) and `p_mpjpe` (loss.py:14-53) on the
device (gast_b200/pipeline.py, csrc/pipeline.cuh).  `mpjpe` is differentiable (forward and backward come
from one kernel pass); `p_mpjpe` takes numpy arrays like the reference (main.py:281-283) or CUDA tensors."""
import numpy as np
import torch

from gast_b200 import pipeline as _P


def mpjpe(predicted, target):
    assert predicted.shape == target.shape
    return _P.mpjpe(predicted, target)


def p_mpjpe(predicted, target):
    assert predicted.shape == target.shape
    as_numpy = not isinstance(predicted, torch.Tensor)
    if as_numpy:
        predicted = torch.as_tensor(np.ascontiguousarray(predicted, dtype=np.float32)).cuda()
        target = torch.as_tensor(np.ascontiguousarray(target, dtype=np.float32)).cuda()
    m = _P.p_mpjpe_per_frame(predicted, target).mean()
    return np.float32(m.item()) if as_numpy else m
