"""Device-side counterpart of a subset of the reference's common/camera.py (same names and signatures; under
`device.` so that the reference module, which has more functions, is not shadowed): `normalize_screen_coordinates`
(:8-12), `image_coordinates` (:15-19), `camera_to_world` (:27-28, one quaternion for all points as
reconstruction.py:204 / gen_skes.py use it).  numpy in -> numpy out (through the device), CUDA tensor in ->
CUDA tensor out."""
import numpy as np
import torch

from gast_b200 import pipeline as _P


def _run(fn, X):
    if isinstance(X, torch.Tensor):
        return fn(X)
    return fn(torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).cuda()).cpu().numpy()


def normalize_screen_coordinates(X, w, h):
    assert X.shape[-1] == 2
    return _run(lambda x: _P.normalize_screen(x, w, h), X)


def image_coordinates(X, w, h):
    assert X.shape[-1] == 2
    return _run(lambda x: _P.normalize_screen(x, w, h, inverse=True), X)


def camera_to_world(X, R, t):
    assert X.shape[-1] == 3
    return _run(lambda x: _P.camera_to_world(x, R, t), X)
