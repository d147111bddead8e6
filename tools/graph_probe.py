"""GPU diagnostic: how much of a bench step is launch gaps?  Times K back-to-back forwards of the bench workload
(a) launched eagerly through the public API and (b) replayed from one CUDA graph of the same forward, and prints the
sum of the per-launch device times next to them."""
import ctypes as C
import os
import sys
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'gast-net-3dposeestimation_b200'))
sys.path.insert(0, REPO)
from gast_b200 import synth  # noqa: E402
from common.skeleton import Skeleton  # noqa: E402
from common.graph_utils import adj_mx_from_skeleton  # noqa: E402
from model.gast_net import SpatioTemporalModel  # noqa: E402


def main():
    B, J, ch, fw, K = 4096, 17, 128, [3, 3, 3], 20
    adj = adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))
    m = SpatioTemporalModel(adj, J, 2, J, fw, channels=ch)
    synth.randomize_module(m, 3)
    m = m.cuda().eval()
    xs = [torch.from_numpy(synth.synth_input(B, 27, J, 2, seed=5 + i)).cuda() for i in range(3)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(5):
            m(xs[0])
        torch.cuda.synchronize()
        e0.record()
        for i in range(K):
            m(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / K
        # one graph per input buffer
        s = torch.cuda.Stream()
        graphs = []
        with torch.cuda.stream(s):
            for i in range(3):
                m(xs[i])
            s.synchronize()
            for i in range(3):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    y = m(xs[i])
                graphs.append(g)
            for g in graphs:
                g.replay()
            s.synchronize()
            e0.record(s)
            for i in range(K):
                graphs[i % 3].replay()
            e1.record(s)
            s.synchronize()
        graph = e0.elapsed_time(e1) / K
    h = m.__dict__['_gast_handles'][('cuda', xs[0].device.index)]
    lib = h.lib
    lib.gast_set_timing(h.h, 1)
    tot = 0.0
    for _ in range(5):
        with torch.no_grad():
            m(xs[0])
        ms = (C.c_float * 512)()
        kinds = (C.c_int32 * 512)()
        n = lib.gast_get_timings(h.h, 512, ms, kinds)
        tot += sum(ms[i] for i in range(n))
    lib.gast_set_timing(h.h, 0)
    print('eager %.3f ms/step, graph replay %.3f ms/step, sum of %d per-launch times %.3f ms' % (eager, graph, n, tot / 5))


if __name__ == '__main__':
    main()
