"""Host-fed lifting of a stream of clip batches: `model_pos(inputs_2d)` of main.evaluate / reconstruction.evaluate
(main.py:309-311, reconstruction.py:160-162) for callers whose keypoints live in host memory.

The reference uploads a batch, runs the forward, and reads the result back one after the other on one stream.
Here the three phases of consecutive batches overlap: a copy stream uploads batch i+1 (pinned host -> one of
`depth` device input slots) and drains result i-1 while the compute stream runs batch i; events order the
hand-overs, nothing synchronises the device until the caller asks.  On a B200 the 15 MB upload of a 4096-clip
batch (~0.6 ms over PCIe) disappears behind the ~8 ms forward.
"""
import torch


class PipelinedLifter(object):
    def __init__(self, model_pos, depth=2):
        self.model = model_pos
        self.device = next(model_pos.parameters()).device
        if self.device.type != 'cuda':
            raise RuntimeError('PipelinedLifter: the model must be on a CUDA device (there is no CPU path)')
        self.depth = int(depth)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._slots = [None] * self.depth

    def _slot(self, k, like):
        s = self._slots[k]
        if s is None or s.shape != like.shape:
            s = torch.empty(like.shape, dtype=torch.float32, device=self.device)
            self._slots[k] = s
        return s

    def run(self, host_batches, host_outputs):
        """host_batches: sequence of pinned float32 (B,T,J,2) tensors; host_outputs: same-length sequence of
        pinned (B,T_out,J,3) tensors that receive the poses.  Returns after everything is ENQUEUED; the compute
        stream (the caller's current stream) waits for the last download, so an event recorded on it afterwards,
        or a synchronize, covers the whole job."""
        n = len(host_batches)
        assert len(host_outputs) == n
        comp = torch.cuda.current_stream(self.device)
        copy = self.copy_stream
        ev_in = [torch.cuda.Event() for _ in range(self.depth)]
        ev_done = [torch.cuda.Event() for _ in range(self.depth)]
        used = [False] * self.depth
        copy.wait_stream(comp)                       # whatever produced the slots / outputs before is done
        with torch.no_grad():
            for i in range(n + 1):
                if i < n:
                    k = i % self.depth
                    x = self._slot(k, host_batches[i])
                    with torch.cuda.stream(copy):
                        if used[k]:
                            copy.wait_event(ev_done[k])          # the forward that read this slot has finished
                        x.copy_(host_batches[i], non_blocking=True)
                        ev_in[k].record(copy)
                    used[k] = True
                if i >= 1:
                    k = (i - 1) % self.depth
                    comp.wait_event(ev_in[k])
                    y = self.model(self._slots[k])
                    ev_done[k].record(comp)
                    with torch.cuda.stream(copy):
                        copy.wait_event(ev_done[k])
                        host_outputs[i - 1].copy_(y, non_blocking=True)
                    y.record_stream(copy)
        comp.wait_stream(copy)
        return host_outputs
