"""GPU parity of the device-side caller steps (csrc/pipeline.cuh through the C ABI and
the same-name modules device/common/{generators,loss,camera}.py, device/tools/mpii_coco_h36m.py) against
outputs of the unmodified reference (tests/golden/pipeline_17.npz) and the numpy oracle.
Bars: bit-exact for index/byte work and the float32 keypoint arithmetic; 1e-6 relative for the float
reductions (stated per test)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import pipeline_oracle as PO
from test_pipeline_oracle import dataset, adam_inputs, LEFT, RIGHT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def G():
    return load_golden('pipeline_17')


def test_chunked_generator_on_device_bit_exact(G):
    from device.common.generators import ChunkedGenerator
    lens, p2, p3, cams = dataset()
    gen = ChunkedGenerator(8, cams, p3, p2, 1, pad=13, causal_shift=0, shuffle=True, random_seed=1234, augment=True,
                           kps_left=LEFT, kps_right=RIGHT, joints_left=LEFT, joints_right=RIGHT)
    assert gen.num_batches == int(G['cg_a_num_batches']) and gen.num_frames() == gen.num_batches * 8
    seen = 0
    for bi, (cam, b3, b2) in enumerate(gen.next_epoch()):
        assert b2.is_cuda and b2.dtype == torch.float32
        if ('cg_a%d_2d' % bi) in G:
            assert np.array_equal(cam.cpu().numpy(), G['cg_a%d_cam' % bi])
            assert np.array_equal(b3.cpu().numpy(), G['cg_a%d_3d' % bi])
            assert np.array_equal(b2.cpu().numpy(), G['cg_a%d_2d' % bi])
            seen += 1
    assert seen == 3
    gen = ChunkedGenerator(5, None, p3, p2, 3, pad=4, causal_shift=4, shuffle=False, augment=False)
    nb = int(G['cg_b_num_batches'])
    batches = list(gen.next_epoch())
    assert len(batches) == nb
    for bi in (0, nb - 1):
        cam, b3, b2 = batches[bi]
        assert cam is None
        assert np.array_equal(b3.cpu().numpy(), G['cg_b%d_3d' % bi])
        assert np.array_equal(b2.cpu().numpy(), G['cg_b%d_2d' % bi])


def test_chunk_gather_large_vs_oracle():
    """Full-size property check: 64 videos, 4096-clip batch of 27-frame windows with flips, against the oracle."""
    from gast_b200 import pipeline as P
    rs = np.random.RandomState(5)
    lens = rs.randint(30, 400, 64)
    p2 = [rs.standard_normal((n, 17, 2)).astype(np.float32) for n in lens]
    p3 = [rs.standard_normal((n, 17, 3)).astype(np.float32) for n in lens]
    pairs = rs.permutation(PO.chunk_pairs(lens, 1, True))[:4096]
    seqs = P.DeviceSequences(p2, p3)
    _, b3, b2 = P.chunk_gather(seqs, pairs, 1, 13, 0, LEFT, RIGHT, LEFT, RIGHT)
    _, r3, r2 = PO.chunk_batch(p2, p3, None, pairs, 1, 13, 0, LEFT, RIGHT, LEFT, RIGHT)
    assert np.array_equal(b2.cpu().numpy(), r2) and np.array_equal(b3.cpu().numpy(), r3)
    with pytest.raises(RuntimeError):
        P.chunk_gather(seqs, pairs[:4], 1, 13, 0, [99], [1], LEFT, RIGHT)


def test_unchunked_generator_on_device():
    from device.common.generators import UnchunkedGenerator
    t = load_golden('tta_17')
    gen = UnchunkedGenerator(None, None, [t['seq']], pad=13, causal_shift=0, augment=True, kps_left=LEFT, kps_right=RIGHT,
                             joints_left=LEFT, joints_right=RIGHT)
    (cam, b3, b2), = list(gen.next_epoch())
    assert cam is None and b3 is None and np.array_equal(b2.cpu().numpy(), t['batch_sym'])
    gen.set_augment(False)
    (_, _, b2), = list(gen.next_epoch())
    assert np.array_equal(b2.cpu().numpy(), t['batch_sym'][:1])


def test_keypoint_formats_bit_exact(G):
    import device.tools.mpii_coco_h36m as K
    for fn, src, key in ((K.coco_h36m, 'k17', 'coco_h36m'), (K.mpii_h36m, 'k16', 'mpii_h36m'),
                         (K.coco_h36m_toe_format, 'k133', 'coco_toe')):
        out, valid = fn(G[src])                                   # numpy in -> numpy out, like the reference
        assert np.array_equal(out, G[key]), key
        assert np.array_equal(valid, G[key + '_valid']), key
        out_t, valid_t = fn(torch.from_numpy(G[src]).cuda())      # device in -> device out
        assert out_t.is_cuda and np.array_equal(out_t.cpu().numpy(), G[key])
        assert np.array_equal(valid_t.cpu().numpy(), G[key + '_valid'])
    rs = np.random.RandomState(3)
    big = rs.uniform(0, 2000, (20000, 17, 2)).astype(np.float32)
    big[::97] = 0
    out, valid = K.coco_h36m(big)
    ref, rvalid = PO.coco_h36m(big)
    assert np.array_equal(out, ref) and np.array_equal(valid, rvalid)
    with pytest.raises(RuntimeError):
        K.coco_h36m(big[:, :16])


def test_camera_functions(G):
    import device.common.camera as Cm
    assert np.array_equal(Cm.normalize_screen_coordinates(G['k17'], w=1920, h=1080), G['norm_screen'])
    assert np.allclose(Cm.image_coordinates(G['norm_screen'], w=1920, h=1080), G['img_coords'], rtol=0, atol=1e-4)
    got = Cm.camera_to_world(G['x3'], R=G['rot'], t=0)
    assert np.abs(got - G['cam2world']).max() < 2e-6               # fp32 cross products, contraction may differ
    dev = Cm.camera_to_world(torch.from_numpy(G['x3']).cuda(), G['rot'], np.array([1.0, 2.0, 3.0], np.float32))
    assert np.abs(dev.cpu().numpy() - (G['cam2world'] + np.array([1, 2, 3], np.float32))).max() < 4e-6


def test_mpjpe_forward_backward(G):
    from device.common.loss import mpjpe
    pred = torch.from_numpy(G['mp_pred']).cuda().requires_grad_(True)
    tgt = torch.from_numpy(G['mp_tgt']).cuda()
    loss = mpjpe(pred, tgt)
    (3.0 * loss).backward()
    assert abs(loss.item() - float(G['mp_loss'])) < 1e-6 * max(1.0, float(G['mp_loss']))
    assert np.abs(pred.grad.cpu().numpy() - 3.0 * G['mp_grad']).max() < 1e-7
    assert pred.grad[0, 0, 3].abs().max().item() == 0.0            # zero distance -> zero gradient, not NaN
    rs = np.random.RandomState(9)                                  # config-3 size and a large one, vs the oracle
    for n in (128 * 17, 1 << 20):
        a = rs.standard_normal((n, 3)).astype(np.float32)
        b = rs.standard_normal((n, 3)).astype(np.float32)
        pa = torch.from_numpy(a).cuda().requires_grad_(True)
        l = mpjpe(pa, torch.from_numpy(b).cuda())
        l.backward()
        rl, rg = PO.mpjpe(a, b)
        assert abs(l.item() - rl) < 1e-6 * rl
        assert np.abs(pa.grad.cpu().numpy() - rg).max() < 1e-6 / n * 10


def test_p_mpjpe(G):
    from device.common.loss import p_mpjpe
    from gast_b200 import pipeline as P
    per = P.p_mpjpe_per_frame(torch.from_numpy(G['pm_pred']).cuda(), torch.from_numpy(G['pm_tgt']).cuda()).cpu().numpy()
    assert np.abs(per - G['pm_per_frame']).max() < 2e-5            # the reference itself runs in float32
    assert np.abs(per - PO.p_mpjpe_per_frame(G['pm_pred'], G['pm_tgt'])).max() < 2e-6
    v = p_mpjpe(G['pm_pred'], G['pm_tgt'])                         # numpy call form of main.py:281-283
    assert abs(float(v) - float(G['pm_value'])) < 2e-5
    rs = np.random.RandomState(4)
    a = rs.standard_normal((5000, 17, 3)).astype(np.float32)
    b = (a @ np.linalg.qr(rs.standard_normal((3, 3)))[0] * 0.7 + 0.05 * rs.standard_normal((5000, 17, 3))).astype(np.float32)
    per = P.p_mpjpe_per_frame(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    assert np.abs(per - PO.p_mpjpe_per_frame(a, b)).max() < 5e-6


def test_fused_adam_amsgrad(G):
    from gast_b200.pipeline import FusedAdam
    ps, grads = adam_inputs()
    params = [torch.nn.Parameter(torch.from_numpy(p.copy()).cuda()) for p in ps]
    opt = FusedAdam(params, lr=1e-3, amsgrad=True)
    for step in range(4):
        for p, g in zip(params, grads[step]):
            p.grad = torch.from_numpy(g.copy()).cuda()
        opt.step()
        for g in opt.param_groups:
            g['lr'] *= 0.95                                        # trainval.py:162-164
    for i, p in enumerate(params):
        ref = G['adam_p%d' % i]
        assert np.abs(p.detach().cpu().numpy() - ref).max() < 2e-6, i
