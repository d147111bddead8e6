"""Small end-to-end workload for compute-sanitizer: one eval forward (tcgen05 core and FFMA core), the fused
forward+mpjpe call, one training step and one pushed frame of a causal stream, each checked against the other path."""
import os
import sys
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'gast-net-3dposeestimation_b200'))
sys.path.insert(0, REPO)
from gast_b200 import engine, synth  # noqa: E402
from gast_b200 import pipeline as P  # noqa: E402
from common.skeleton import Skeleton  # noqa: E402
from common.graph_utils import adj_mx_from_skeleton  # noqa: E402
from model.gast_net import SpatioTemporalModel, SpatioTemporalModelOptimized1f  # noqa: E402


def main():
    J, fw, ch, B = 17, [3, 3, 3], 32, 9
    adj = adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))
    m = SpatioTemporalModel(adj, J, 2, J, fw, channels=ch)
    synth.randomize_module(m, 3)
    m = m.cuda().eval()
    x = torch.from_numpy(synth.synth_input(B, 27, J, 2, seed=5)).cuda()
    t = torch.from_numpy(synth.synth_target(B, J, seed=6)).cuda()
    with torch.no_grad():
        y_tc = m(x)
        engine.set_gemm_core(1)
        y_ff = m(x)
        engine.set_gemm_core(0)
        y2, loss = engine.run_model_mpjpe(m, x, t)
    torch.cuda.synchronize()
    print('eval: tc vs ffma max diff %.2e, fused loss %.6f vs %.6f' % ((y_tc - y_ff).abs().max().item(), loss.item(),
                                                                      P.mpjpe(y_tc, t).item()))
    assert torch.equal(y2, y_tc)
    mt = SpatioTemporalModelOptimized1f(adj, J, 2, J, fw, dropout=0.1, channels=ch)
    synth.randomize_module(mt, 4)
    mt = mt.cuda().train()
    opt = torch.optim.SGD(mt.parameters(), lr=1e-3)
    for _ in range(2):
        opt.zero_grad()
        l = P.mpjpe(mt(x), t)
        l.backward()
        opt.step()
    torch.cuda.synchronize()
    print('train: loss %.6f' % l.item())


if __name__ == '__main__':
    main()
