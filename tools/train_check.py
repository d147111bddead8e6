"""GPU diagnostic: per-parameter gradient error of the training path vs the torch-CPU port (autograd)."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'gast-net-3dposeestimation_b200')); sys.path.insert(0, REPO)
from gast_b200 import synth
from oracle import gast_torch_ref as TR, gast_oracle as O
from model.gast_net import SpatioTemporalModelOptimized1f
from common.skeleton import Skeleton
from common.graph_utils import adj_mx_from_skeleton

J, fw, ch, B = 17, [3, 3, 3], int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 16
adj = adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))
masks = tuple(torch.from_numpy(a) for a in O.local_masks(O.adj_from_parents(synth.skeleton_parents(J))))
m = SpatioTemporalModelOptimized1f(adj, J, 2, J, fw, dropout=0.0, channels=ch)
synth.randomize_module(m, 5)
p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in m.state_dict().items()}
x = torch.from_numpy(synth.synth_input(B, 27, J, 2, seed=3)); tgt = torch.from_numpy(synth.synth_target(B, J, seed=4))
y_ref = TR.forward(x, p, masks, fw, strided=True, training=True, stats={})
TR.mpjpe(y_ref, tgt).backward()
# double-precision oracle to separate rounding noise from bugs
pd = {k: (v.detach().double().requires_grad_(v.requires_grad) if v.dtype.is_floating_point else v.detach().clone()) for k, v in p.items()}
yd = TR.forward(x.double(), pd, masks, fw, strided=True, training=True, stats={})
TR.mpjpe(yd, tgt.double()).backward()
m = m.cuda().train()
xc = x.cuda()
y = m(xc)
torch.mean(torch.norm(y - tgt.cuda(), dim=3)).backward()
print('forward max err', (y.detach().cpu() - y_ref.detach()).abs().max().item())
rows = []
for k, prm in m.named_parameters():
    g, gr, gd = prm.grad.cpu(), p[k].grad, pd[k].grad
    mx = gd.abs().max().item() + 1e-30
    rows.append((k, (g.double() - gd).abs().max().item() / mx, (gr.double() - gd).abs().max().item() / mx, mx))
rows.sort(key=lambda t: -t[1])
with open(os.path.join(REPO, 'gpurun_out', 'train_grad_report.txt'), 'w') as f:
    for r in rows:
        f.write('%-75s cuda_vs_fp64 %.3e   cpu32_vs_fp64 %.3e   max|g| %.3e\n' % r)
for r in rows[:25]:
    print('%-75s cuda_vs_fp64 %.3e   cpu32_vs_fp64 %.3e   max|g| %.3e' % r)
