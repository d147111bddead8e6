"""GPU diagnostic: device time of every launch of one bench-workload forward, in launch order
(CUDA events inside the library: gast_set_timing / gast_get_timings)."""
import ctypes as C
import os
import sys
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'gast-net-3dposeestimation_b200'))
sys.path.insert(0, REPO)
from gast_b200 import engine, synth  # noqa: E402
from common.skeleton import Skeleton  # noqa: E402
from common.graph_utils import adj_mx_from_skeleton  # noqa: E402
from model.gast_net import SpatioTemporalModel  # noqa: E402


def main():
    # usage: launch_times.py [clips [joints [channels [filter,widths]]]]
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    J = int(sys.argv[2]) if len(sys.argv) > 2 else 17
    ch = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    fw = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else [3, 3, 3]
    T = 1
    for v in fw:
        T *= v
    adj = adj_mx_from_skeleton(Skeleton(synth.skeleton_parents(J), [], []))
    m = SpatioTemporalModel(adj, J, 2, J, fw, channels=ch)
    synth.randomize_module(m, 3)
    m = m.cuda().eval()
    x = torch.from_numpy(synth.synth_input(B, T, J, 2, seed=5)).cuda()
    with torch.no_grad():
        for _ in range(3):
            m(x)
    h = m.__dict__['_gast_handles'][('cuda', x.device.index)]
    lib = h.lib
    lib.gast_set_timing(h.h, 1)
    reps, tot = 5, None
    for _ in range(reps):
        with torch.no_grad():
            m(x)
        ms = (C.c_float * 512)()
        kinds = (C.c_int32 * 512)()
        n = lib.gast_get_timings(h.h, 512, ms, kinds)
        cur = [ms[i] for i in range(n)]
        tot = cur if tot is None else [a + b for a, b in zip(tot, cur)]
    lib.gast_set_timing(h.h, 0)
    for i in range(n):
        print('%2d %-16s %.4f ms' % (i, engine.LAUNCH_KINDS[kinds[i]], tot[i] / reps))
    print('sum %.3f ms' % (sum(tot) / reps))


if __name__ == '__main__':
    main()
