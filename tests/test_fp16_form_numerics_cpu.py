"""CPU emulation of the operand arithmetic of the tcgen05 core (csrc/gemm_tc.cuh), independent of the GPU:

* PREC 0 (K < 256 and GAST_TC_F16=0): x_hi = tf32(x), x_lo = bf16(x - x_hi), D = A_hi.B_hi + A_lo.bf16(B_hi) + bf16(A_hi).B_lo
* PREC 2 (default for K >= 256): x_hi = fp16(x) saturating, x_lo = fp16(x - x_hi), weights as 2^8 W,
  D = (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo) * 2^-8

with exact (float64) accumulation, so that only the operand rounding is measured: the numbers DESIGN.md §4.1 quotes for
the two forms (the GPU measurement incl. the tensor core's accumulation is profiles/r02_z4_f16_probe.txt), the
degradation outside fp16's comfortable range, and the host-side guard of tc_prepare_weights."""
import numpy as np
import pytest
import torch


def tf32(x):
    u = np.asarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16(x):
    return torch.from_numpy(np.asarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def f16(x):
    return np.clip(x, -65504, 65504).astype(np.float16).astype(np.float32)


def prec0(A, W):
    Ah, Wh = tf32(A), tf32(W)
    Al, Wl = bf16(A - Ah), bf16(W - Wh)
    d = lambda a, b: a.astype(np.float64) @ b.astype(np.float64).T
    return d(Ah, Wh) + d(Al, bf16(Wh)) + d(bf16(Ah), Wl)


def prec2(A, W):
    Ws = (W * np.float32(256)).astype(np.float32)
    Ah, Wh = f16(A), f16(Ws)
    Al, Wl = f16(A - Ah), f16(Ws - Wh)
    d = lambda a, b: a.astype(np.float64) @ b.astype(np.float64).T
    return (d(Ah, Wh) + d(Al, Wh) + d(Ah, Wl)) / 256.0


def weights_fit_fp16(W):
    """mirror of the guard in tc_prepare_weights: 2^8 max|W| < 60000 (NaN does not fit)"""
    m = np.abs(W).max()
    return bool(m * 256.0 < 60000.0)


def rel_rms(y, ref):
    return float(np.sqrt(((y - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))


def case(ascale, wscale, K=512, seed=0):
    rng = np.random.default_rng(seed)
    A = np.maximum(rng.standard_normal((192, K)) * ascale, 0).astype(np.float32)        # post-ReLU activations
    W = (rng.standard_normal((96, K)) * wscale).astype(np.float32)
    return A, W, A.astype(np.float64) @ W.astype(np.float64).T


@pytest.mark.parametrize('ascale,wscale', [(1, 0.05), (30, 0.05), (1, 0.9), (3000, 0.05), (1, 1e-3), (0.05, 0.05)])
def test_fp16_form_is_fp32_grade_in_the_working_range(ascale, wscale):
    A, W, ref = case(ascale, wscale)
    e0, e2 = rel_rms(prec0(A, W), ref), rel_rms(prec2(A, W), ref)
    assert e0 < 1e-6                                  # tf32 + bf16 corrections: ~6.5e-7
    assert e2 < 5e-7                                  # 22 significant bits per operand: ~1e-7 (3.5e-7 at |x| ~ 0.05)
    assert weights_fit_fp16(W)


def test_fp16_form_degrades_gracefully_outside_it():
    # tiny activations / tiny weights: remainders fall into fp16's subnormals (absolute step 6e-8): still far below 1e-4
    A, W, ref = case(1e-3, 0.05)
    assert rel_rms(prec2(A, W), ref) < 5e-5
    A, W, ref = case(1, 3e-5)
    assert rel_rms(prec2(A, W), ref) < 1e-5
    # the other form keeps fp32's range
    assert rel_rms(prec0(A, W), ref) < 1e-6


def test_weight_guard_rule():
    rng = np.random.default_rng(1)
    W = rng.standard_normal((64, 256)).astype(np.float32)
    assert weights_fit_fp16(W * 10)                   # 2^8 * ~40 = 1e4
    assert not weights_fit_fp16(W * 300)              # 2^8 * ~1200 = 3e5 > 65504: would saturate, keeps tf32 + bf16
    Wn = W.copy()
    Wn[3, 5] = np.nan
    assert not weights_fit_fp16(Wn)
    # what the guard prevents: the saturated split is useless
    A, Wb, ref = case(1, 300.0)
    assert rel_rms(prec2(A, Wb), ref) > 1e-2
    assert rel_rms(prec0(A, Wb), ref) < 1e-6
