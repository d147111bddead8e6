"""Shared helpers for the reference-generated training fixtures (tests/golden/train_*.npz, written by
tests/golden/make_golden.py:train_case from the UNMODIFIED reference in train() mode, driven like main.train(),
main.py:219-239, with optim.Adam(lr=1e-3, amsgrad=True), trainval.py:78)."""
import numpy as np
import torch

from conftest import load_golden
from gast_b200 import synth


def adj_t(J):
    from common.skeleton import Skeleton
    from common.graph_utils import adj_mx_from_skeleton
    return adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))


def build_module(meta):
    """our drop-in module shell with the fixture's weights (synth is keyed on state_dict names, so these are
    the weights the reference model had)"""
    from model.gast_net import SpatioTemporalModel, SpatioTemporalModelOptimized1f
    J = meta['J']
    cls = SpatioTemporalModel if meta['dilated'] else SpatioTemporalModelOptimized1f
    m = cls(adj_t(J), J, 2, J, meta['filter_widths'], dropout=0.0, channels=meta['channels'])
    synth.randomize_module(m, meta['seed'])
    assert [k for k, _ in m.named_parameters()] == meta['names']
    return m


def batch(meta, step):
    x = torch.from_numpy(synth.synth_input(meta['B'], meta['T'], meta['J'], 2, seed=100 + step))
    tgt = torch.from_numpy(synth.synth_target(meta['B'], meta['J'], seed=200 + step))
    tgt[:, :, 0] = 0                                    # main.py:225
    return x, tgt


def digest(t, idx):
    f = t.detach().reshape(-1).double().cpu()
    return np.array([float(f.norm()), float(f.sum())], np.float64), f[torch.from_numpy(idx)].numpy().astype(np.float32)


def check_step(g, step, y, loss, grads, y_tol, loss_rtol, ent_rtol, norm_rtol, report=None):
    """compare one step of an implementation with the fixture.  grads: name -> tensor.
    ent_rtol: tolerance on the stored large-|g| entries relative to the largest of them;
    norm_rtol: tolerance on ||g||.  Every violation is collected (and reported) before the assert."""
    bad = []
    yr = g['y%d' % step]
    scale = max(1.0, float(np.abs(yr).max()))
    ey = float(np.abs(y - yr).max())
    if not ey < y_tol * scale:
        bad.append(('y', step, ey))
    lr = float(g['loss%d' % step])
    el = abs(loss - lr) / abs(lr)
    if not el < loss_rtol:
        bad.append(('loss', step, loss, lr))
    # gradients that are zero by construction (init_bn.bias: expand_bn removes any constant it adds) are pure
    # rounding noise in ANY implementation, the reference included: they only have to stay noise-sized
    gmax = max(float(g['gsum%d/%s' % (step, k)][0]) for k in g['meta']['names'])
    worst = (0.0, None)
    for k in g['meta']['names']:
        ns_ref, ent_ref = g['gsum%d/%s' % (step, k)], g['gent%d/%s' % (step, k)]
        ns, ent = digest(grads[k], g['idx/' + k])
        if ns_ref[0] < 1e-5 * gmax:
            if not ns[0] < 1e-4 * gmax:
                bad.append(('noise-level gradient grew', step, k, ns[0]))
            continue
        en = abs(ns[0] - ns_ref[0]) / ns_ref[0]
        ee = float(np.abs(ent - ent_ref).max() / np.abs(ent_ref).max())
        if max(en, ee) > worst[0]:
            worst = (max(en, ee), k)
        if not en < norm_rtol:
            bad.append(('grad norm', step, k, ns[0], ns_ref[0]))
        if not ee < ent_rtol:
            bad.append(('grad entries', step, k, ee))
    if report is not None:
        report.append(('step %d: max|dy| %.3g  loss rel %.3g  worst grad fingerprint %.3g (%s)  violations %d'
                       % (step, ey, el, worst[0], worst[1], len(bad))))
        report.extend('   ' + repr(b) for b in bad[:20])
    return bad


__all__ = ['load_golden', 'build_module', 'batch', 'digest', 'check_step', 'adj_t']
