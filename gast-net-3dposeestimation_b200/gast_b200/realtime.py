"""Frame-by-frame causal lifting of many concurrent keypoint streams (first step).

The reference's real-time model is a causal `SpatioTemporalModelOptimized1f` fed the last `receptive_field`
frames of one stream (`gen_skes.py:43-69`: `causal=True`; the window ends at the current frame).  `CausalStream`
keeps those windows for `n_streams` streams in ONE device tensor used as a ring (no per-frame reallocation, no host
round trip), pads the start of a stream by replicating its first frame exactly like
`UnchunkedGenerator(pad, causal_shift=pad)` does (`common/generators.py:217-221`), and lifts all streams with one
forward per pushed frame.  Every push recomputes the receptive field (O(rf) per frame, as the reference does); the
O(1)-per-frame variant with per-stage activation rings is the next step of N4.
"""
import torch


class CausalStream(object):
    def __init__(self, model_pos, n_streams, device=None):
        self.model = model_pos
        self.rf = int(model_pos.receptive_field())
        if device is None:
            device = next(model_pos.parameters()).device
        self.device = torch.device(device)
        self.n = int(n_streams)
        self.J = int(model_pos.num_joints_in)
        self.F = int(model_pos.in_features)
        # ring[:, (head + k) % rf] is the k-th oldest frame of the window; `window` is its unrolled copy
        self.ring = torch.zeros((self.n, self.rf, self.J, self.F), dtype=torch.float32, device=self.device)
        self.window = torch.empty_like(self.ring)
        self.head = 0
        self.started = torch.zeros(self.n, dtype=torch.bool, device=self.device)

    def reset(self, streams=None):
        """Forget the history of all (or the listed) streams: their next frame starts a new sequence."""
        if streams is None:
            self.started.zero_()
        else:
            self.started[torch.as_tensor(streams, device=self.device)] = False

    def push(self, frame):
        """frame (n_streams, J, in_features): the newest 2D keypoints of every stream (normalised screen
        coordinates).  Returns the 3D pose of that frame, (n_streams, J, 3)."""
        frame = torch.as_tensor(frame, dtype=torch.float32, device=self.device)
        assert frame.shape == (self.n, self.J, self.F), frame.shape
        fresh = ~self.started
        if bool(fresh.any()):
            # a stream's first frame fills its whole window: the edge padding of generators.py:217-221
            self.ring[fresh] = frame[fresh].unsqueeze(1)
            self.started |= fresh
        self.ring[:, self.head] = frame                      # overwrite the oldest slot
        self.head = (self.head + 1) % self.rf
        k = self.rf - self.head                              # unroll: oldest frame first, newest last
        self.window[:, :k] = self.ring[:, self.head:]
        self.window[:, k:] = self.ring[:, :self.head]
        with torch.no_grad():
            return self.model(self.window)[:, -1]
