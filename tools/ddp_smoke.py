"""torchrun smoke of the data-parallel training step: ranks end with identical parameters."""
import os, sys
import torch, torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'gast-net-3dposeestimation_b200')); sys.path.insert(0, REPO)
from gast_b200 import synth
from gast_b200.trainer import DataParallelTrainer
from gast_b200.dist import shard_range
from model.gast_net import SpatioTemporalModelOptimized1f
from common.skeleton import Skeleton
from common.graph_utils import adj_mx_from_skeleton

rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
J = 17
m = SpatioTemporalModelOptimized1f(adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], [])), J, 2, J, [3, 3, 3],
                                   dropout=0.0, channels=128)
synth.randomize_module(m, 1)
m = m.cuda()
tr = DataParallelTrainer(m, lambda ps: torch.optim.Adam(ps, lr=1e-3, amsgrad=True))
B = 128
x = torch.from_numpy(synth.synth_input(B, 27, J, 2, seed=1)); y = torch.from_numpy(synth.synth_target(B, J, seed=2))
s, e = shard_range(B, rank, world)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for step in range(6):
    if step == 2:
        torch.cuda.synchronize(); dist.barrier(); ev0.record()
    loss = tr.step(x[s:e].cuda(), y[s:e].cuda())
ev1.record(); torch.cuda.synchronize()
flat = torch.cat([p.detach().flatten() for p in m.parameters()])
chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
allc = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
if rank == 0:
    print('ddp_smoke: world %d, local batch %d, loss %.6f, params identical across ranks: %s, %.2f ms/step (fwd+bwd+allreduce+Adam)'
          % (world, e - s, loss.item(), same, ev0.elapsed_time(ev1) / 4))
assert same
dist.destroy_process_group()
