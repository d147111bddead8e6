"""GPU probe of the tcgen05 kind::tf32 accumulate numerics (through gast_debug_gemm).
Dumps small cases for offline bit-level emulation and prints bias statistics vs fp64."""
import os
import sys
import ctypes as C
import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'gast-net-3dposeestimation_b200'))
from gast_b200 import _lib  # noqa: E402

lib = _lib.load()
OUT = os.path.join(REPO, 'gpurun_out')
os.makedirs(OUT, exist_ok=True)


def tf32_round(x):
    """round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def gemm(A, W, core, mode):
    a = torch.from_numpy(A).cuda().contiguous()
    w = torch.from_numpy(W).cuda().contiguous()
    M, K = A.shape
    N = W.shape[0]
    o = torch.empty((M, N), dtype=torch.float32, device='cuda')
    rc = lib.gast_debug_gemm(a.data_ptr(), w.data_ptr(), o.data_ptr(), M, N, K, core, mode,
                             torch.cuda.current_stream().cuda_stream)
    if rc:
        raise RuntimeError(_lib.last_error())
    return o.cpu().numpy()


def stats(name, D, ref):
    err = (D.astype(np.float64) - ref)
    rel = err / np.maximum(np.abs(ref), 1e-30)
    shrink = (np.abs(D.astype(np.float64)) - np.abs(ref)) / np.maximum(np.abs(ref), 1e-30)
    big = np.abs(ref) > 0.1 * np.abs(ref).mean()
    print('%-46s max|err| %.3e  rms rel %.3e  mean shrink(|D|-|ref|)/|ref| %.3e' %
          (name, np.abs(err).max(), np.sqrt((rel[big] ** 2).mean()), shrink[big].mean()), flush=True)


def main():
    rs = np.random.RandomState(0)
    dump = {}
    # --- 1. single / few MMAs, tf32-exact operands, hi.hi only ------------------------------
    for nk in (8, 16, 32, 64):
        K = max(32, nk)
        A = np.zeros((128, K), np.float32)
        W = np.zeros((128, K), np.float32)
        A[:, :nk] = tf32_round(rs.standard_normal((128, nk)) * np.exp2(rs.randint(-6, 6, (128, nk))))
        W[:, :nk] = tf32_round(rs.standard_normal((128, nk)))
        D = gemm(A, W, 0, 0)   # tf32-exact operands: lo parts are exactly 0, so this is hi.hi only
        ref = A.astype(np.float64) @ W.astype(np.float64).T
        stats('hi.hi only, %d k (tf32-exact, mixed exp)' % nk, D, ref)
        dump['A%d' % nk] = A
        dump['W%d' % nk] = W
        dump['D%d' % nk] = D
    np.savez_compressed(os.path.join(OUT, 'tc_probe_small.npz'), **dump)
    # --- 2. statistics at real K --------------------------------------------------------------
    for K in (128, 512, 1536):
        for kind in ('randsign', 'positive'):
            A = rs.standard_normal((256, K)).astype(np.float32)
            W = (rs.standard_normal((128, K)) / np.sqrt(K)).astype(np.float32)
            if kind == 'positive':
                A, W = np.abs(A), np.abs(W)
            ref = A.astype(np.float64) @ W.astype(np.float64).T
            stats('K=%4d %-8s FFMA fp32' % (K, kind), gemm(A, W, 1, 0), ref)
            stats('K=%4d %-8s tcgen05 3xTF32' % (K, kind), gemm(A, W, 0, 0), ref)
            Ah, Wh = tf32_round(A), tf32_round(W)
            refh = Ah.astype(np.float64) @ Wh.astype(np.float64).T
            stats('K=%4d %-8s tcgen05 hi.hi vs exact(hi.hi)' % (K, kind), gemm(Ah, Wh, 0, 0), refh)


if __name__ == '__main__':
    main()
