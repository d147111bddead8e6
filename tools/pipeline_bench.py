"""GPU measurement of the device-side caller steps: device time per call (CUDA events,
after warm-up, L2 flushed by rotating buffers where it matters) and achieved HBM GB/s on ALGORITHMIC bytes
(inputs read once + outputs written once).  Prints one JSON object; run under gpurun."""
import json
import os
import sys
import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'gast-net-3dposeestimation_b200'))
from gast_b200 import pipeline as P  # noqa: E402

LEFT, RIGHT = [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    peak = 6487.4
    try:
        peak = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    out = {'hbm_peak_gbs': peak, 'kernels': {}}

    def rec(name, ms, nbytes, note):
        out['kernels'][name] = {'ms': ms, 'algorithmic_bytes': nbytes, 'gbs': nbytes / ms / 1e6,
                                'frac_of_hbm_peak': nbytes / ms / 1e6 / peak, 'note': note}

    rs = np.random.RandomState(0)
    # N1: one 4096-clip training batch of 27-frame windows out of 600 videos x 3000 frames (245 MB of 2D poses)
    lens = [3000] * 600
    p2 = [rs.standard_normal((n, 17, 2)).astype(np.float32) for n in lens]
    p3 = [rs.standard_normal((n, 17, 3)).astype(np.float32) for n in lens]
    seqs = P.DeviceSequences(p2, p3)
    pairs = np.stack([rs.randint(0, 600, 4096), rs.randint(0, 3000, 4096), np.zeros(4096, np.int64), rs.randint(0, 2, 4096)], 1)
    pairs[:, 2] = pairs[:, 1] + 1
    ms = timed(lambda: P.chunk_gather(seqs, pairs, 1, 13, 0, LEFT, RIGHT, LEFT, RIGHT))
    nb = 4096 * (27 * 17 * 2 + 17 * 3) * 4 * 2
    rec('chunk_gather_4096x27f', ms, nb, 'includes the host->device copy of the 64 KB pair list and two launches; '
        '%.0f k clips/s' % (4096 / ms))
    # N3
    k = torch.from_numpy(rs.uniform(0, 1000, (1 << 20, 17, 2)).astype(np.float32)).cuda()
    ms = timed(lambda: P.keypoints_convert(k, P.KPT_COCO_H36M))
    rec('coco_h36m_1Mframes', ms, (1 << 20) * (17 * 2 * 4 * 2 + 4), 'one thread per frame (136-byte rows: strided access)')
    ms = timed(lambda: P.normalize_screen(k, 1920, 1080))
    rec('normalize_screen_17.8Mpoints', ms, k.numel() * 8, '')
    x3 = torch.from_numpy(rs.standard_normal((1 << 20, 17, 3)).astype(np.float32)).cuda()
    ms = timed(lambda: P.camera_to_world(x3, [0.14, -0.15, -0.755, 0.622], 0))
    rec('camera_to_world_17.8Mpoints', ms, x3.numel() * 8, '')
    # N2
    a = torch.from_numpy(rs.standard_normal((4096, 1, 17, 3)).astype(np.float32)).cuda().requires_grad_(True)
    b = torch.from_numpy(rs.standard_normal((4096, 1, 17, 3)).astype(np.float32)).cuda()
    ms = timed(lambda: P.mpjpe(a, b))
    rec('mpjpe_fwd_bwd_4096clips', ms, a.numel() * 12, 'two launches (partial sums + final), latency-bound at this size')
    big_a = torch.from_numpy(rs.standard_normal((1 << 22, 3)).astype(np.float32)).cuda().requires_grad_(True)
    big_b = torch.from_numpy(rs.standard_normal((1 << 22, 3)).astype(np.float32)).cuda()
    ms = timed(lambda: P.mpjpe(big_a, big_b))
    rec('mpjpe_fwd_bwd_4Mpoints', ms, big_a.numel() * 12, '')
    pa = torch.from_numpy(rs.standard_normal((1 << 18, 17, 3)).astype(np.float32)).cuda()
    pb = torch.from_numpy(rs.standard_normal((1 << 18, 17, 3)).astype(np.float32)).cuda()
    ms = timed(lambda: P.p_mpjpe_per_frame(pa, pb))
    rec('p_mpjpe_262kframes', ms, pa.numel() * 8 + (1 << 18) * 4, 'one warp per frame, fp64 3x3 Jacobi SVD: compute-bound (%.1f M frames/s)' % ((1 << 18) / ms / 1e3))
    shapes = [(6915984,)]                       # the 27f/17j/128ch parameter count
    ps = [torch.nn.Parameter(torch.zeros(sh, device='cuda')) for sh in shapes]
    for p_ in ps:
        p_.grad = torch.randn_like(p_)
    opt = P.FusedAdam(ps, lr=1e-3, amsgrad=True)
    ms = timed(lambda: opt.step())
    rec('adam_amsgrad_6.9Mparams', ms, 6915984 * 4 * (2 + 3 + 1 + 3), 'reads p,g,m,v,vmax; writes p,m,v,vmax')
    tp = [torch.nn.Parameter(torch.zeros(sh, device='cuda')) for sh in shapes]
    for p_ in tp:
        p_.grad = torch.randn_like(p_)
    topt = torch.optim.Adam(tp, lr=1e-3, amsgrad=True)
    out['torch_adam_amsgrad_6.9Mparams_ms'] = timed(lambda: topt.step())
    print(json.dumps(out))


if __name__ == '__main__':
    main()
