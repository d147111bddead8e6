"""GPU: frame-by-frame causal lifting (gast_b200.realtime.CausalStream) equals the whole-sequence
causal forward the reference would run (UnchunkedGenerator(pad, causal_shift=pad) + dilated SpatioTemporalModel,
main.py:299-320).  Runs last: it only composes kernels the other GPU tests have already checked."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from gast_b200 import synth
from gast_b200.realtime import CausalStream
from test_gpu_parity import build_model, TOL

pytestmark = pytest.mark.gpu


def test_causal_stream_equals_full_sequence_forward():
    g1f = load_golden('model_17_333_c16_causal_1f_T27')          # causal Optimized1f, as gen_skes.load_model_realtime builds
    gfull = load_golden('model_17_333_c16_causal_full_T30')      # the same network as a dilated causal SpatioTemporalModel
    m1f = build_model(g1f['meta'])
    mfull = build_model(gfull['meta'])
    mfull.load_state_dict(m1f.state_dict())                      # identical key lists (main.py:252 relies on it)
    rf = m1f.receptive_field()
    T, n = 40, 3
    seqs = np.stack([synth.synth_input(1, T, 17, 2, seed=50 + s)[0] for s in range(n)])      # (n, T, 17, 2)
    cs = CausalStream(m1f, n_streams=n)
    got = torch.stack([cs.push(torch.from_numpy(seqs[:, t]).cuda()) for t in range(T)], dim=1)   # (n, T, 17, 3)
    padded = np.pad(seqs, ((0, 0), (rf - 1, 0), (0, 0), (0, 0)), 'edge')                      # pad + causal_shift on the left
    with torch.no_grad():
        want = mfull(torch.from_numpy(padded).cuda())
    assert want.shape == got.shape == (n, T, 17, 3)
    assert (got - want).abs().max().item() < TOL
