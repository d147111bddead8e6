"""GPU diagnostic: run each sub-module on the tcgen05 core and on the FFMA core, report the
difference, and dump arrays to gpurun_out/ when they disagree (for offline analysis)."""
import os
import sys
import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root
sys.path.insert(0, os.path.join(REPO, 'gast-net-3dposeestimation_b200'))
sys.path.insert(0, REPO)
from gast_b200 import engine, synth  # noqa: E402
from common.skeleton import Skeleton  # noqa: E402
from common.graph_utils import adj_mx_from_skeleton  # noqa: E402
from model.gast_net import GraphAttentionBlock, SpatioTemporalModel  # noqa: E402
from model.local_attention import LocalGraph, SemCHGraphConv, local_adjacencies  # noqa: E402
from model.global_attention import MultiGlobalGraph, GlobalGraph  # noqa: E402

OUT = os.path.join(REPO, 'gpurun_out')
os.makedirs(OUT, exist_ok=True)


def adj(J):
    return adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))


def both(name, mod, x, tc_expected=True):
    mod = mod.cuda().eval()
    outs = []
    for core in (1, 0):
        engine.set_gemm_core(core)
        with torch.no_grad():
            y = mod(x)
        torch.cuda.synchronize()
        outs.append(y.float().cpu().numpy().copy())
    engine.set_gemm_core(0)
    err = np.abs(outs[0] - outs[1]).max()
    ref = np.abs(outs[0]).max()
    print('%-28s max|ffma-tc| = %.3e   (max|y| = %.3e)  %s' % (name, err, ref, 'OK' if err < 2e-5 * max(ref, 1) else 'MISMATCH'),
          flush=True)
    if not err < 2e-5 * max(ref, 1):
        np.save(os.path.join(OUT, name + '_ffma.npy'), outs[0])
        np.save(os.path.join(OUT, name + '_tc.npy'), outs[1])
    return err


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    torch.manual_seed(0)
    J = 17
    for C in (32, 128):
        F = 23  # frames: 4 tiles of 7 frames, last one ragged
        x = torch.from_numpy((0.7 * np.random.RandomState(C).standard_normal((1, F, J, C))).astype(np.float32)).cuda()
        _, con = local_adjacencies(adj(J))
        if which in ('all', 'semch'):
            m = SemCHGraphConv(C, C, con); synth.randomize_module(m, 1)
            both('semch_C%d' % C, m, x)
        if which in ('all', 'ghead'):
            m = GlobalGraph(adj(J), C, C // 4); synth.randomize_module(m, 2)
            both('globalhead_C%d' % C, m, x.reshape(F, J, C).permute(0, 2, 1).contiguous())
        if which in ('all', 'local'):
            m = LocalGraph(adj(J), C, C, 0.05); synth.randomize_module(m, 3)
            both('local_C%d' % C, m, x)
        if which in ('all', 'mglobal'):
            m = MultiGlobalGraph(adj(J), C, C // 4, 0.05); synth.randomize_module(m, 4)
            both('mglobal_C%d' % C, m, x)
        if which in ('all', 'block'):
            m = GraphAttentionBlock(adj(J), C, C, 0.05); synth.randomize_module(m, 5)
            both('block_C%d' % C, m, x.permute(0, 3, 1, 2).contiguous())
    if which in ('all', 'model'):
        for (Jm, ch, B, T) in ((17, 32, 5, 27), (17, 32, 2, 45), (19, 64, 9, 27), (17, 128, 300, 27)):
            m = SpatioTemporalModel(adj(Jm), Jm, 2, Jm, [3, 3, 3], channels=ch); synth.randomize_module(m, 6)
            xm = torch.from_numpy(synth.synth_input(B, T, Jm, 2, 7)).cuda()
            both('model_J%d_C%d_B%d_T%d' % (Jm, ch, B, T), m, xm)


if __name__ == '__main__':
    main()
