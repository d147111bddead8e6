"""CPU: the ring-buffer windowing of gast_b200.realtime.WindowStream (the generic O(rf)-per-frame driver) against the reference's causal padding
(UnchunkedGenerator(pad, causal_shift=pad): common/generators.py:217-221), with a stub in place of the network
(the network itself has no CPU path)."""
import numpy as np
import torch

from gast_b200.realtime import WindowStream


class WindowProbe(torch.nn.Module):
    """Stands in for a causal Optimized1f: 'lifts' a window to a weighted sum over its frames, so that any
    mistake in the order or the padding of the window changes the result."""

    def __init__(self, rf, J=17):
        super().__init__()
        self.w = torch.nn.Parameter(torch.arange(1, rf + 1, dtype=torch.float32) ** 1.5, requires_grad=False)
        self.rf, self.num_joints_in, self.in_features = rf, J, 2

    def receptive_field(self):
        return self.rf

    def forward(self, x):                                    # (n, rf, J, 2) -> (n, 1, J, 3)
        assert x.shape[1] == self.rf
        s = (x * self.w.view(1, -1, 1, 1)).sum(1, keepdim=True)
        return torch.cat([s, s[..., :1]], dim=-1)


def reference_windows(seq, rf):
    """what the reference feeds a causal model for every frame of a sequence: pad = rf-1 frames on the left
    (pad + causal_shift with causal_shift = pad), none on the right, edge mode."""
    padded = np.pad(seq, ((rf - 1, 0), (0, 0), (0, 0)), 'edge')
    return np.stack([padded[t:t + rf] for t in range(seq.shape[0])])


def test_causal_stream_windows_match_reference_padding():
    rs = np.random.RandomState(0)
    rf, J, T = 27, 17, 70
    seqs = rs.standard_normal((3, T, J, 2)).astype(np.float32)
    probe = WindowProbe(rf, J)
    cs = WindowStream(probe, n_streams=3, device='cpu')
    outs = []
    for t in range(T):
        outs.append(cs.push(seqs[:, t]).numpy().copy())
        if t == 40:
            cs.reset([1])                                    # stream 1 restarts at t = 41
    outs = np.stack(outs, 1)                                 # (3, T, J, 3)
    for s in range(3):
        if s == 1:
            exp = np.concatenate([reference_windows(seqs[s, :41], rf), reference_windows(seqs[s, 41:], rf)])
        else:
            exp = reference_windows(seqs[s], rf)
        want = probe(torch.from_numpy(exp))[:, -1].numpy()   # (T, J, 3)
        assert np.allclose(outs[s], want, rtol=1e-6, atol=1e-5), s
