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
