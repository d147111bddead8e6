"""Training path (BASELINE configs[2]) on the GPU against the torch-CPU port of
the reference under autograd: train-mode forward (batch-stat BN, running-stat update), every
parameter gradient, and a 3-step Adam(amsgrad) loop with loss / state_dict match (dropout 0)."""
import numpy as np
import pytest
import torch

from gast_b200 import synth

pytestmark = pytest.mark.gpu


def _adj(J):
    from common.skeleton import Skeleton
    from common.graph_utils import adj_mx_from_skeleton
    return adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))


def _masks(J):
    from oracle import gast_oracle as O
    return tuple(torch.from_numpy(a) for a in O.local_masks(O.adj_from_parents(synth.skeleton_parents(J))))


def _oracle_step(p, x, target, fw, J, stats):
    from oracle import gast_torch_ref as TR
    y = TR.forward(x, p, _masks(J), fw, strided=True, training=True, stats=stats)
    return y, TR.mpjpe(y, target)


@pytest.mark.parametrize('J,fw,ch,B', [(17, [3, 3, 3], 32, 6), (17, [3, 3, 3], 128, 16), (19, [3, 3], 32, 5)])
def test_train_forward_and_gradients(J, fw, ch, B):
    from model.gast_net import SpatioTemporalModelOptimized1f
    T = int(np.prod(fw))
    m = SpatioTemporalModelOptimized1f(_adj(J), J, 2, J, fw, dropout=0.0, channels=ch)
    synth.randomize_module(m, 5)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.from_numpy(synth.synth_input(B, T, J, 2, seed=3))
    tgt = torch.from_numpy(synth.synth_target(B, J, seed=4))
    # oracle: autograd through the torch-CPU port, in fp32 (forward / loss / running statistics) and in
    # fp64 (gradients: fp32 gradients of this net carry ~1e-2 max-normalised noise from ReLU decisions
    # of near-zero BatchNorm outputs -- the fp32 CPU port is itself that far from fp64, see
    # profiles/r01_train_grad_report.txt -- so the gradient oracle has to be the fp64 one)
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd0.items()}
    stats = {}
    y_ref, loss_ref = _oracle_step(p, x, tgt, fw, J, stats)
    loss_ref.backward()
    pd = {k: (v.detach().double().requires_grad_(v.requires_grad) if v.dtype.is_floating_point else v.clone())
          for k, v in p.items()}
    _, loss_d = _oracle_step(pd, x.double(), tgt.double(), fw, J, {})
    loss_d.backward()
    # CUDA path
    m = m.cuda().train()
    y = m(x.cuda())
    loss = torch.mean(torch.norm(y - tgt.cuda(), dim=3))
    loss.backward()
    scale = float(y_ref.detach().abs().max())
    loss_ref = loss_ref.detach()
    assert (y.detach().cpu() - y_ref.detach()).abs().max().item() < 5e-5 * max(scale, 1.0)
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
    # running statistics updated like nn.BatchNorm2d(momentum=0.1)
    sd1 = m.state_dict()
    for k, v in stats.items():
        assert (sd1[k].cpu() - v).abs().max().item() < 1e-5 * max(1.0, float(v.abs().max())), k
    assert int(sd1['init_bn.num_batches_tracked']) == 1
    # every parameter gradient
    bad = []
    for k, prm in m.named_parameters():
        assert prm.grad is not None, k
        g, gd = prm.grad.cpu().double(), pd[k].grad
        if gd.abs().max().item() < 1e-7:       # zero by construction (a bias in front of a batch-stat BN)
            assert g.abs().max().item() < 1e-5, k
            continue
        l2 = ((g - gd).norm() / gd.norm()).item()
        mx = ((g - gd).abs().max() / gd.abs().max()).item()
        g32 = p[k].grad.double()                      # the reference port's own fp32 gradient
        l2_ref = ((g32 - gd).norm() / gd.norm()).item()
        mx_ref = ((g32 - gd).abs().max() / gd.abs().max()).item()
        # as accurate as the reference's fp32 arithmetic (x4 slack), or better than 1e-2 / 3e-2
        if not (l2 < max(1e-2, 4 * l2_ref) and mx < max(3e-2, 4 * mx_ref)):
            bad.append((k, round(l2, 5), round(l2_ref, 5), round(mx, 5), round(mx_ref, 5)))
    assert not bad, sorted(bad, key=lambda t: -t[1])[:12]


def _run_steps(make_opt, nsteps, B=32, seed=9):
    from model.gast_net import SpatioTemporalModelOptimized1f
    J, fw, ch = 17, [3, 3, 3], 128
    m = SpatioTemporalModelOptimized1f(_adj(J), J, 2, J, fw, dropout=0.0, channels=ch)
    synth.randomize_module(m, seed)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in sd0.items()}
    opt_ref = make_opt([v for k, v in p.items() if v.requires_grad])
    m = m.cuda().train()
    opt = make_opt(list(m.parameters()))
    stats, losses, g_ref = {}, [], None
    for step in range(nsteps):
        x = torch.from_numpy(synth.synth_input(B, 27, J, 2, seed=100 + step))
        tgt = torch.from_numpy(synth.synth_target(B, J, seed=200 + step))
        opt_ref.zero_grad()
        _, loss_ref = _oracle_step(p, x, tgt, fw, J, stats)
        loss_ref.backward()
        g_ref = {k: v.grad.clone() for k, v in p.items() if v.requires_grad}
        opt_ref.step()
        opt.zero_grad()
        y = m(x.cuda())
        loss = torch.mean(torch.norm(y - tgt.cuda(), dim=3))
        loss.backward()
        opt.step()
        losses.append((loss.item(), loss_ref.item()))
    return m, p, stats, losses, g_ref


def test_three_sgd_steps_match_reference_port():
    """BASELINE configs[2] (Optimized1f [3,3,3]/128ch, here b=32 for CI time): forward + backward +
    optimiser step, three times; the loss must track the reference port to 1e-4 relative and the
    weights / running statistics must agree afterwards.  SGD(momentum) is used for the multi-step
    check because Adam's first steps are ~lr*sign(g): parameters whose gradient is rounding noise
    move by +-lr at random in ANY two implementations (CPU vs GPU reference included)."""
    m, p, stats, losses, _ = _run_steps(lambda ps: torch.optim.SGD(ps, lr=1e-2, momentum=0.9), 3)
    for step, (a, b) in enumerate(losses):
        assert abs(a - b) < 1e-4 * abs(b), (step, a, b)
    sd = m.state_dict()
    for k, v in p.items():
        # init_bn.bias has an identically zero gradient (expand_bn removes any constant it adds), so
        # its drift is pure rounding noise in any implementation and it cannot influence the output
        if 'num_batches' in k or 'running' in k or k == 'init_bn.bias':
            continue
        d = ((sd[k].cpu() - v.detach()).norm() / (v.detach().norm() + 1e-12)).item()
        assert d < 1e-3, (k, d)
    for k, v in stats.items():
        assert (sd[k].cpu() - v).abs().max().item() < 1e-4 * max(1.0, float(v.abs().max())), k
    m.eval()                                   # eval after training sees the new weights and statistics
    with torch.no_grad():
        ye = m(torch.from_numpy(synth.synth_input(4, 27, 17, 2, seed=1)).cuda())
    assert torch.isfinite(ye).all()


def test_one_adam_amsgrad_step_matches_reference_port():
    """trainval.py:78 uses Adam(amsgrad=True): one step, parameters compared where the gradient is
    well above the fp32 gradient noise (|g| > 0.1 max|g| of the tensor:
    Adam's first step is lr*sign(g), so a noise-level entry moves by +-lr at random)."""
    m, p, stats, losses, g_ref = _run_steps(lambda ps: torch.optim.Adam(ps, lr=1e-3, amsgrad=True), 1)
    assert abs(losses[0][0] - losses[0][1]) < 1e-5 * abs(losses[0][1])
    sd = m.state_dict()
    checked = 0
    for k, g in g_ref.items():
        if k == 'init_bn.bias':               # identically zero gradient: rounding noise only
            continue
        sel = g.abs() > 0.1 * g.abs().max()      # fp32 gradients carry up to ~1e-2 max-normalised noise
        if sel.any():
            d = (sd[k].cpu() - p[k].detach())[sel].abs().max().item()
            assert d < 2e-5, (k, d)
            checked += int(sel.sum())
    assert checked > 100000


def test_dropout_and_mode_errors():
    from model.gast_net import SpatioTemporalModel, SpatioTemporalModelOptimized1f
    m = SpatioTemporalModelOptimized1f(_adj(17), 17, 2, 17, [3, 3, 3], dropout=0.25, channels=32).cuda().train()
    x = torch.from_numpy(synth.synth_input(4, 27, 17, 2)).cuda()
    torch.manual_seed(0)
    y1 = m(x)
    torch.manual_seed(0)
    y2 = m(x)
    torch.manual_seed(1)
    y3 = m(x)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3)      # dropout stream follows torch's seed
    y1.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    d = SpatioTemporalModel(_adj(17), 17, 2, 17, [3, 3, 3], channels=32).cuda().train()
    with pytest.raises(RuntimeError, match='Optimized1f'):
        d(x)
