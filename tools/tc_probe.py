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
    rc = lib.gast_debug_gemm(a.data_ptr(), w.data_ptr(), o.data_ptr(), M, N, K, core, mode, 0, None,
                             torch.cuda.current_stream().cuda_stream)
    if rc:
        raise RuntimeError(_lib.last_error())
    return o.cpu().numpy()


def perf():
    """time the tcgen05 kernel and its experiment variants on the lifting path's GEMM shapes"""
    rs = np.random.RandomState(1)
    for (M, N, K) in ((674176, 256, 384), (674176, 128, 128), (224768, 512, 768), (75008, 1024, 1536)):
        a = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).cuda()
        w = torch.from_numpy(rs.standard_normal((N, K)).astype(np.float32)).cuda()
        o = torch.empty((M, N), dtype=torch.float32, device='cuda')
        mma_bound = (2.0 if b"bf16-corr" in lib.gast_version() else 3.0) * (-(-M // 128) * 128) * (-(-N // 128) * 128) * K / (2048.0 * 148 * 1.9e9) * 1e3
        line = 'M=%d N=%d K=%d  MMA-bound %.3f ms |' % (M, N, K, mma_bound)
        for core, mode, name in ((1, 0, 'ffma'), (0, 0, 'tc'), (0, 1, 'noflush'), (0, 2, 'noAload'), (0, 3, 'mainonly'),
                                 (0, 4, 'noSTTM'), (0, 5, 'mma+B only'), (0, 7, 'corr-only'), (0, 8, 'mma-only(no A/B traffic)'), (0, 9, 'mma+barriers only')):
            ms = C.c_float(0)
            rc = lib.gast_debug_gemm(a.data_ptr(), w.data_ptr(), o.data_ptr(), M, N, K, core, mode, 5, C.byref(ms),
                                     torch.cuda.current_stream().cuda_stream)
            line += ' %s %.3f' % (name, ms.value) if rc == 0 else ' %s ERR(%s)' % (name, _lib.last_error()[:40])
        print(line, flush=True)
        # clock64() attribution (DBG 6): cycles per chunk spent waiting, per role
        ms = C.c_float(0)
        rc = lib.gast_debug_gemm(a.data_ptr(), w.data_ptr(), o.data_ptr(), M, N, K, 0, 6, 0, C.byref(ms),
                                 torch.cuda.current_stream().cuda_stream)
        if rc == 0:
            d = o.view(-1).view(torch.int64)[:148 * 32].reshape(148, 32).double().cpu().numpy()
            nA = d[:, 0].mean(); nM = d[:, 16].mean(); nE = max(d[:, 24].mean(), 1)
            print('   per chunk [cycles]: A-prod global-data wait+STS %.0f, LDS+split %.0f, refill issue %.0f, wait a_empty %.0f, STTM+arrive %.0f, loop total %.0f | B-prod wait b_empty %.0f |'
                  ' MMA wait main_empty %.0f, a_full %.0f, b_full %.0f, issue+commit %.0f | epi wait main_full %.0f of %.0f per group'
                  % (d[:, 4].mean() / nA, d[:, 5].mean() / nA, d[:, 6].mean() / nA, d[:, 1].mean() / nA, d[:, 2].mean() / nA, d[:, 3].mean() / nA, d[:, 8].mean() / nM,
                     d[:, 17].mean() / nM, d[:, 18].mean() / nM, d[:, 19].mean() / nM, d[:, 20].mean() / nM,
                     d[:, 25].mean() / nE, d[:, 26].mean() / nE), flush=True)
            nT = max(d[:, 27].mean(), 1)
            print('   epilogue per tile [cycles]: total %.0f | first-group flush %.0f | wait corr %.0f | corr load+add %.0f | => bias/act/store ~%.0f' % (d[:, 26].mean() / nT, d[:, 30].mean() / nT, d[:, 28].mean() / nT, d[:, 29].mean() / nT, (d[:, 26].mean() - d[:, 25].mean() - d[:, 30].mean() - d[:, 28].mean() - d[:, 29].mean()) / nT), flush=True)
        else:
            print('   dbg6 failed:', _lib.last_error())


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
            stats('K=%4d %-8s tcgen05 tf32+bf16corr' % (K, kind), gemm(A, W, 0, 0), ref)
            stats('K=%4d %-8s tcgen05 3xTF32 (training)' % (K, kind), gemm(A, W, 2, 0), ref)
            Ah, Wh = tf32_round(A), tf32_round(W)
            refh = Ah.astype(np.float64) @ Wh.astype(np.float64).T
            stats('K=%4d %-8s tcgen05 hi.hi vs exact(hi.hi)' % (K, kind), gemm(Ah, Wh, 0, 0), refh)


def train_perf():
    """FFMA vs tcgen05 3xTF32 on the GEMM shapes of a b = 128 training step (forward / dgrad)"""
    rs = np.random.RandomState(2)
    for (M, N, K) in ((19584, 512, 128), (19584, 256, 384), (6528, 1024, 256), (6528, 512, 768), (6528, 256, 768),
                      (2176, 2048, 512), (2176, 1024, 1536), (2176, 512, 1536), (2176, 1536, 1024)):
        a = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).cuda()
        w = torch.from_numpy(rs.standard_normal((N, K)).astype(np.float32)).cuda()
        o = torch.empty((M, N), dtype=torch.float32, device='cuda')
        line = 'M=%d N=%d K=%d |' % (M, N, K)
        for core, name in ((1, 'ffma'), (2, 'tc 3xTF32'), (0, 'tc tf32+bf16')):
            ms = C.c_float(0)
            rc = lib.gast_debug_gemm(a.data_ptr(), w.data_ptr(), o.data_ptr(), M, N, K, core, 0, 10, C.byref(ms),
                                     torch.cuda.current_stream().cuda_stream)
            line += ' %s %.4f ms (%.0f TF/s)' % (name, ms.value, 2e-9 * M * N * K / max(ms.value, 1e-6)) if rc == 0 else ' %s ERR' % name
        print(line, flush=True)


def cg_perf():
    """full kernel / MMAs without operand traffic / MMAs + barrier hand-shakes only, for the cta_group selected by GAST_TC_CG"""
    rs = np.random.RandomState(1)
    print('GAST_TC_CG=%s' % os.environ.get('GAST_TC_CG', 'default'))
    for (M, N, K) in ((75008, 1024, 1536), (224768, 512, 768), (674176, 128, 128)):
        a = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).cuda()
        w = torch.from_numpy(rs.standard_normal((N, K)).astype(np.float32)).cuda()
        o = torch.empty((M, N), dtype=torch.float32, device='cuda')
        line = 'M=%d N=%d K=%d |' % (M, N, K)
        for mode, name in ((0, 'tc'), (4, 'noSTTM'), (5, 'mma+B only'), (8, 'mma-only(no A/B traffic)'), (9, 'mma+barriers only')):
            ms = C.c_float(0)
            rc = lib.gast_debug_gemm(a.data_ptr(), w.data_ptr(), o.data_ptr(), M, N, K, 0, mode, 5, C.byref(ms),
                                     torch.cuda.current_stream().cuda_stream)
            line += ' %s %.3f' % (name, ms.value) if rc == 0 else ' %s ERR(%s)' % (name, _lib.last_error()[:40])
        print(line, flush=True)


def f16_probe():
    """fp16-hi arithmetic (gast_debug_gemm core 3: A_h16.B_h16 + A_lo.B_h16 + A_h16.B_lo, kind::f16 MMAs on fp16 operands, weights as 2^8 W)
    next to tf32 + bf16 corrections (core 0; the remainders are fp16 too: mixed fp16 / bf16 operands of one MMA are illegal) and the FFMA core: error against fp64 and time on the path's GEMM shapes"""
    rs = np.random.RandomState(0)

    def stats(name, o, ref):
        d = o.astype(np.float64) - ref
        sc = np.sqrt((ref ** 2).mean())
        print('%-52s rel rms %.2e  max %.2e  bias %.2e' % (name, np.sqrt((d ** 2).mean()) / sc, np.abs(d).max() / sc, d.mean() / sc),
              flush=True)
    for K in (128, 384, 1536):
        for kind in ('randsign', 'positive', 'relu_x_small_w'):
            A = rs.standard_normal((256, K)).astype(np.float32)
            W = (rs.standard_normal((128, K)) / np.sqrt(K)).astype(np.float32)
            if kind == 'positive':
                A, W = np.abs(A), np.abs(W)
            if kind == 'relu_x_small_w':
                A, W = np.maximum(A, 0) * 30, W * 1e-3
            ref = A.astype(np.float64) @ W.astype(np.float64).T
            for core, name in ((1, 'FFMA fp32'), (0, 'tcgen05 tf32 + bf16 corr'), (3, 'tcgen05 fp16 hi + fp16 lo')):
                stats('K=%4d %-14s %s' % (K, kind, name), gemm(A, W, core, 0), ref)
    rs = np.random.RandomState(1)
    for (M, N, K) in ((75008, 1024, 1536), (224768, 512, 768), (674176, 128, 128), (626688, 256, 384)):
        a = torch.from_numpy(rs.standard_normal((M, K)).astype(np.float32)).cuda()
        w = torch.from_numpy(rs.standard_normal((N, K)).astype(np.float32)).cuda()
        o = torch.empty((M, N), dtype=torch.float32, device='cuda')
        line = 'M=%d N=%d K=%d |' % (M, N, K)
        for core, name in ((0, 'tf32+bf16'), (3, 'fp16hi')):
            ms = C.c_float(0)
            rc = lib.gast_debug_gemm(a.data_ptr(), w.data_ptr(), o.data_ptr(), M, N, K, core, 0, 5, C.byref(ms),
                                     torch.cuda.current_stream().cuda_stream)
            line += ' %s %.4f ms (%.0f TF/s)' % (name, ms.value, 2e-9 * M * N * K / max(ms.value, 1e-6)) if rc == 0 else ' %s ERR(%s)' % (name, _lib.last_error()[:60])
        print(line, flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--f16':
        f16_probe()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == '--cg':
        cg_perf()
    elif len(sys.argv) > 1 and sys.argv[1] == '--train':
        train_perf()
    elif len(sys.argv) > 1 and sys.argv[1] == '--perf':
        perf()
    else:
        main()
