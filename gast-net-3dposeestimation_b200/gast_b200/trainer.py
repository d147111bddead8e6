"""Data-parallel training step for the lifting network (BASELINE configs[2]).

One process per GPU; each rank runs forward + backward on its own clips (BatchNorm statistics
are per rank, like the replicas of the reference's nn.DataParallel, trainval.py:56-61), then ONE
NCCL all-reduce over a flat fp32 gradient buffer (27.66 MB at 27f/17j/128ch) averages the
gradients, and every rank applies the same optimiser step.  With world == 1 the all-reduce is the
identity, so the loss matches the single-device reference exactly (dropout 0).
"""
import torch
import torch.distributed as dist

from .dist import FlatGradBuffer


def mpjpe(predicted, target):
    """common/loss.py:5-11"""
    assert predicted.shape == target.shape
    return torch.mean(torch.norm(predicted - target, dim=len(target.shape) - 1))


class DataParallelTrainer(object):
    def __init__(self, model, optimizer_factory, group=None):
        self.model = model
        self.group = group
        self.flat = FlatGradBuffer(model.parameters())        # .grad of every parameter aliases one buffer
        # one backward per step into buffers this trainer owns and zeroes: the library writes the gradients in place
        # (engine._TrainFn.backward) instead of handing 166 tensors to autograd's per-parameter accumulation
        model.__dict__['_gast_direct_grads'] = True
        self.opt = optimizer_factory(model.parameters())
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def step(self, inputs_2d, inputs_3d):
        """main.train()'s inner loop (main.py:219-239) for this rank's shard of the batch."""
        self.model.train()
        inputs_3d = inputs_3d.clone()
        inputs_3d[:, :, 0] = 0                                 # main.py:225
        self.flat.zero_()
        pred = self.model(inputs_2d)
        loss = mpjpe(pred, inputs_3d)
        loss.backward()                                        # accumulates into the (zeroed) flat buffer
        self.flat.all_reduce_mean(self.group)                  # the only collective of the path
        self.opt.step()
        return loss.detach()


class GraphedTrainer(DataParallelTrainer):
    """The same step with the forward, the loss and the backward replayed from ONE CUDA graph.

    A b = 128 step is ~700 small launches whose host-side enqueue (14.8 ms) costs as much as their device time
    (16 ms, `profiles/r02_train_step.txt`): the step is launch-bound on the host.  Here the launches of
    `flat.zero_() -> model(x) -> mpjpe -> backward` are captured once (static input / target buffers, the workspace
    and the gradient views come from the graph's private pool) and replayed per step; the dropout masks stay fresh
    through the library's device-side dropout counter (`gast_set_dropout_state`), BatchNorm's running statistics
    and `num_batches_tracked` are updated by the captured kernels.  The gradient all-reduce and the one-launch
    optimiser step run after the replay, outside the graph."""

    def __init__(self, model, optimizer_factory, batch_shape, target_shape, group=None, warmup=3):
        super().__init__(model, optimizer_factory, group)
        dev = self.flat.flat.device
        self.x = torch.zeros(batch_shape, dtype=torch.float32, device=dev)
        self.y = torch.zeros(target_shape, dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.model.__dict__['_gast_dropout_state'] = torch.zeros(1, dtype=torch.int64, device=dev)
        self.model.train()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # allocator, attribute opt-ins, handle creation: outside capture
                self._fwd_bwd()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._fwd_bwd()

    def _fwd_bwd(self):
        self.flat.zero_()
        pred = self.model(self.x)
        loss = mpjpe(pred, self.y)
        loss.backward()
        self.loss.copy_(loss.detach())

    def step(self, inputs_2d, inputs_3d):
        self.x.copy_(inputs_2d, non_blocking=True)
        self.y.copy_(inputs_3d, non_blocking=True)
        self.y[:, :, 0] = 0                              # main.py:225
        self.graph.replay()
        self.flat.all_reduce_mean(self.group)
        self.opt.step()
        return self.loss
