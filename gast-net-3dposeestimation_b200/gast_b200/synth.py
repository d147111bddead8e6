"""Deterministic synthetic weights and inputs (no checkpoints/datasets exist offline).

The reference ships no trained weights (README.md:78-87 are download links), so parity is
pinned on *seeded, randomised* state: every state_dict entry is filled from a numpy
`RandomState` keyed on (seed, crc32(key)), which is stable across machines and independent
of torch's RNG and of key order.  Defaults (`e`=1, `C_k`=0, biases 0, BN stats 0/1) would
hide bugs, so they are all randomised.
"""
import zlib
import numpy as np
import torch


def _rs(seed, key):
    return np.random.RandomState((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 32))


def synth_value(key, shape, seed):
    """float32 numpy array for state_dict entry `key` of `shape`."""
    rs = _rs(seed, key)
    shape = tuple(int(s) for s in shape)
    leaf = key.split('.')[-1]
    parent = key.split('.')[-2] if '.' in key else ''
    is_bn = parent.startswith('bn') or parent.endswith('_bn') or (
        parent.isdigit() and 'layers_bn' in key)
    if leaf == 'num_batches_tracked':
        return np.zeros(shape, dtype=np.int64)
    if leaf == 'running_mean':
        return (0.1 * rs.standard_normal(shape)).astype(np.float32)
    if leaf == 'running_var':
        return rs.uniform(0.5, 1.5, shape).astype(np.float32)
    if is_bn and leaf == 'weight':
        return rs.uniform(0.5, 1.5, shape).astype(np.float32)
    if is_bn and leaf == 'bias':
        return (0.1 * rs.standard_normal(shape)).astype(np.float32)
    if leaf == 'e':
        return (1.0 + 0.5 * rs.standard_normal(shape)).astype(np.float32)
    if leaf == 'C_k':
        return (0.05 * rs.standard_normal(shape)).astype(np.float32)
    if leaf == 'bias':  # conv biases of g / theta / phi
        return (0.3 * rs.standard_normal(shape)).astype(np.float32)
    if leaf == 'W':  # SemCH (2, Cin, Cout)
        std = np.sqrt(2.0 / (shape[1] + shape[2]))
        return (std * rs.standard_normal(shape)).astype(np.float32)
    if leaf == 'weight':  # conv weights (Cout, Cin, k...) ; He-like so activations stay O(1)
        fan_in = int(np.prod(shape[1:]))
        if 'concat_project' in key:
            std = 1.0 / np.sqrt(fan_in)
        elif key.startswith('shrink'):
            std = 1.0 / np.sqrt(fan_in)
        else:
            std = np.sqrt(1.5 / fan_in)
        return (std * rs.standard_normal(shape)).astype(np.float32)
    raise KeyError('synth: unknown state_dict entry %r' % key)


def randomize_module(module, seed=1):
    """Overwrite every parameter/buffer of `module` in place with synth_value()."""
    sd = module.state_dict()
    with torch.no_grad():
        for k, v in sd.items():
            val = synth_value(k, v.shape, seed)
            v.copy_(torch.from_numpy(val).to(v.dtype))
    return module


def synth_state(keys_shapes, seed=1):
    """{key: numpy array} for an ordered list of (key, shape)."""
    return {k: synth_value(k, s, seed) for k, s in keys_shapes}


def synth_input(B, T, J, F=2, seed=1234):
    """Screen-normalised-like keypoints: clip(0.5*N(0,1), -1, 1), float32 (B,T,J,F)."""
    rs = np.random.RandomState(int(seed) % (2 ** 32))
    x = 0.5 * rs.standard_normal((B, T, J, F))
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def synth_target(B, J, seed=4321):
    """Synthetic 3D ground truth (B,1,J,3), root joint zeroed (main.py:225)."""
    rs = np.random.RandomState(int(seed) % (2 ** 32))
    y = (0.5 * rs.standard_normal((B, 1, J, 3))).astype(np.float32)
    y[:, :, 0] = 0
    return y


H36M_PARENTS_17 = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15]
H36M_PARENTS_19 = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 10, 13, 14, 10, 16, 17]
HUMANEVA_PARENTS_15 = [-1, 0, 1, 2, 3, 1, 5, 6, 0, 8, 9, 0, 11, 12, 1]


def skeleton_parents(J):
    """Parent arrays the reference uses (reconstruction.py:87,95; gast_net.py:265)."""
    if J == 17:
        return list(H36M_PARENTS_17)
    if J == 19:
        return list(H36M_PARENTS_19)
    if J == 15:
        return list(HUMANEVA_PARENTS_15)
    if J == 16:
        # 17-joint H36M skeleton with joint 9 (neck/nose) removed, stacked-hourglass layout
        p = list(H36M_PARENTS_17)
        del p[9]
        return [q if q < 9 else q - 1 for q in [8 if q == 9 else q for q in p]]
    raise KeyError(J)
