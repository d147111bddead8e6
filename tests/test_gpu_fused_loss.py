"""GPU: forward + mpjpe in one call (gast_forward_mpjpe: the loss is taken in the shrink kernel's epilogue) against the
oracle forward + the reference formula of common/loss.py:5-11, and against the two-call path of the library."""
import numpy as np
import pytest
import torch

from gast_b200 import engine, synth
from gast_b200 import pipeline as P

pytestmark = pytest.mark.gpu


def _model(J, fw, ch, seed=3):
    from model.gast_net import SpatioTemporalModel
    from common.skeleton import Skeleton
    from common.graph_utils import adj_mx_from_skeleton
    adj = adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))
    m = SpatioTemporalModel(adj, J, 2, J, fw, channels=ch)
    synth.randomize_module(m, seed)
    return m


@pytest.mark.parametrize('J,fw,ch,B', [(17, [3, 3, 3], 32, 37), (19, [3, 3], 16, 5), (17, [3, 3, 3], 128, 300)])
def test_forward_mpjpe_matches_two_calls_and_oracle(J, fw, ch, B):
    from oracle import gast_oracle as O
    m = _model(J, fw, ch)
    T = int(np.prod(fw))
    x = synth.synth_input(B, T, J, 2, seed=11)
    tgt = synth.synth_target(B, J, seed=12)
    mc = m.cuda().eval()
    xc, tc = torch.from_numpy(x).cuda(), torch.from_numpy(tgt).cuda()
    with torch.no_grad():
        y2 = mc(xc)
        l2 = P.mpjpe(y2, tc)
        y1, l1 = engine.run_model_mpjpe(mc, xc, tc)
    # same forward, bit for bit (the fused shrink keeps shrink_kernel's FMA order)
    assert torch.equal(y1, y2)
    # same loss as the separate mpjpe kernel (both sum in double; the float result may differ in the last bit)
    assert abs(l1.item() - l2.item()) <= 2e-7 * max(1.0, abs(l2.item()))
    # against the oracle forward + the reference formula (float64), small case only
    if ch <= 32:
        p = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        ref = O.forward(x, p, O.adj_from_parents(synth.skeleton_parents(J)), fw)
        lref = np.linalg.norm(ref.astype(np.float64) - tgt, axis=-1).mean()
        assert abs(l1.item() - lref) <= 1e-4, (l1.item(), lref)
    # run to run identical (fixed-order reduction behind the atomic ticket), and the ticket is reset
    with torch.no_grad():
        _, l3 = engine.run_model_mpjpe(mc, xc, tc)
    assert l3.item() == l1.item()


def test_forward_mpjpe_rejects_bad_arguments():
    m = _model(17, [3, 3, 3], 16).cuda().eval()
    x = torch.from_numpy(synth.synth_input(2, 27, 17, 2, seed=1)).cuda()
    with pytest.raises(engine.GastError):
        engine.run_model_mpjpe(m, x, torch.zeros(2, 1, 17, 2, device='cuda'))     # wrong target shape
    with pytest.raises(engine.GastError):
        engine.run_model_mpjpe(m, x, torch.zeros(2, 1, 17, 3))                     # CPU target: no CPU path
