"""CPU emulation of the index arithmetic of the streaming kernels in csrc/kernels_hbm.cuh (the kernels themselves are
checked on the GPU by the parity suite): every output element is produced exactly once from the right inputs.

* global_mix_tile_kernel: item -> (frame slot, joint group, channel quad), the transposed attention slab
  att_s[frame][head][j][slot(i)] with slot(i) = (i / IW) * 4 + i % IW, host-side choice of IW / frames per group / threads;
* expand_rows_kernel: staged gather index ((b*T + t*stride + kk)*J + j)*Fin + i against the convolution it implements;
* rowdot8_stream_kernel: chunk -> rows partition over warps of a persistent grid.
Reference formulas: global_attention.py:74-80, gast_net.py:163-164, global_attention.py:60-72."""
import numpy as np
import pytest


def host_mix_plan(J, Ng, max_threads=384):
    """mirror of the launch code in gast_api.cu (run_global, GAST_MIX_MODE=2)"""
    GV = Ng // 4
    IW = 3 if ((J + 2) // 3) * 3 < ((J + 3) // 4) * 4 else 4
    NIG = (J + IW - 1) // IW
    per_frame = NIG * GV
    fpb = max(1, max_threads // per_frame)
    total = fpb * per_frame
    passes = (total + max_threads - 1) // max_threads
    nt = (((total + passes - 1) // passes) + 31) // 32 * 32
    return IW, fpb, nt


def emulate_mix_tile(F, J, heads, Cg, IW, fpb, NT, rng):
    H2, Ng = 2 * heads, heads * Cg
    G = rng.standard_normal((F * J, Ng)).astype(np.float32)
    ab = rng.standard_normal((F * J, H2)).astype(np.float32)
    ck = rng.standard_normal((heads, J, J)).astype(np.float32)
    Yref = np.zeros_like(G)
    for f in range(F):
        for h in range(heads):
            a, b = ab[f * J:(f + 1) * J, 2 * h], ab[f * J:(f + 1) * J, 2 * h + 1]
            s = a[:, None] + b[None, :]
            s = np.where(s >= 0, s, 0.2 * s)
            e = np.exp(s - s.max(1, keepdims=True))
            att = e / e.sum(1, keepdims=True) + ck[h]
            Yref[f * J:(f + 1) * J, h * Cg:(h + 1) * Cg] = att @ G[f * J:(f + 1) * J, h * Cg:(h + 1) * Cg]
    NIG = (J + IW - 1) // IW
    JP, GV = NIG * 4, Ng // 4
    Y = np.full_like(G, np.nan)
    cnt = np.zeros(G.shape, int)
    for grp in range((F + fpb - 1) // fpb):
        f0 = grp * fpb
        nf = min(fpb, F - f0)
        att_s = np.zeros(fpb * heads * J * JP, np.float32)
        abq = ab[f0 * J:(f0 + nf) * J].reshape(-1)
        for e in range(fpb * heads * J):
            i, hh, fr = e % J, (e // J) % heads, e // (J * heads)
            if fr >= nf:
                continue
            abr = abq[fr * J * H2:]
            v = np.array([abr[i * H2 + 2 * hh] + abr[j * H2 + 2 * hh + 1] for j in range(J)], np.float32)
            v = np.where(v >= 0, v, np.float32(0.2) * v)
            v = np.exp(v - v.max())
            inv = 1 / v.sum()
            base = ((fr * heads + hh) * J) * JP + (i // IW) * 4 + (i % IW)
            for j in range(J):
                att_s[base + j * JP] = v[j] * inv + ck[hh, i, j]
        items = fpb * NIG * GV
        for tid in range(NT):
            for w in range(tid, items, NT):
                cq, t = w % GV, w // GV
                ig, fs = t % NIG, t // NIG
                if fs >= nf:
                    continue
                c = cq * 4
                h = c // Cg
                o = np.zeros((IW, 4), np.float32)
                for j in range(J):
                    g = G[(f0 + fs) * J + j, c:c + 4]
                    a4 = att_s[((fs * heads + h) * J) * JP + ig * 4 + j * JP:][:4]
                    for u in range(IW):
                        o[u] += a4[u] * g
                for u in range(IW):
                    if ig * IW + u < J:
                        r = (f0 + fs) * J + ig * IW + u
                        Y[r, c:c + 4] = o[u]
                        cnt[r, c:c + 4] += 1
    return Y, Yref, cnt


@pytest.mark.parametrize('F,J,heads,Cg', [(5, 17, 4, 8), (5, 19, 4, 8), (3, 16, 4, 16), (7, 15, 2, 8), (2, 17, 4, 32)])
def test_mix_tile_items_cover_every_output_once(F, J, heads, Cg):
    IW, fpb, NT = host_mix_plan(J, heads * Cg)
    assert IW == {15: 3, 16: 4, 17: 3, 19: 4}[J]
    assert NT % 32 == 0 and NT <= 384
    Y, Yref, cnt = emulate_mix_tile(F, J, heads, Cg, IW, fpb, NT, np.random.default_rng(J))
    assert (cnt == 1).all()
    assert np.abs(Y - Yref).max() < 2e-5


def test_mix_plan_for_bench_shapes():
    # 27f/17j/128ch: blocks at C = 128, 256, 512 ; 81f/64ch first block ; 19 joints
    assert host_mix_plan(17, 128) == (3, 2, 384)
    assert host_mix_plan(17, 256) == (3, 1, 384)
    assert host_mix_plan(17, 512) == (3, 1, 384)      # 768 items in two passes of 384 threads
    assert host_mix_plan(17, 64) == (3, 4, 384)
    assert host_mix_plan(19, 128) == (4, 2, 320)


@pytest.mark.parametrize('taps,Fin,stride,T,J', [(3, 2, 3, 27, 17), (3, 2, 1, 9, 5), (5, 2, 5, 25, 4)])
def test_expand_rows_gather_index(taps, Fin, stride, T, J):
    rng = np.random.default_rng(0)
    B = 3
    T0 = (T - taps) // stride + 1
    x = rng.standard_normal((B, T, J, Fin)).astype(np.float32)
    flat = x.reshape(-1)
    rows = B * T0 * J
    KF = taps * Fin
    for row in rng.integers(0, rows, 64):
        f, j = divmod(int(row), J)
        b, t = divmod(f, T0)
        got = [flat[((b * T + t * stride + k // Fin) * J + j) * Fin + k % Fin] for k in range(KF)]
        want = x[b, t * stride:t * stride + taps, j, :].reshape(-1)
        assert np.array_equal(np.array(got, np.float32), want)


@pytest.mark.parametrize('rows,K,grid', [(1000, 128, 3), (37, 256, 2), (5, 512, 4), (626688 // 64, 128, 7)])
def test_rowdot_stream_chunks_partition_rows(rows, K, grid):
    cb = max(2048, 2 * K * 4)
    rpc = cb // (4 * K)
    nchunks = (rows + rpc - 1) // rpc
    GW = grid * 8
    seen = np.zeros(rows, int)
    for gw in range(GW):
        ci = gw
        while ci < nchunks:
            r0 = ci * rpc
            nr = min(rpc, rows - r0)
            assert nr * K * 4 % 16 == 0 and nr >= 1
            for ra in range(0, nr, 2):
                seen[r0 + ra] += 1
                if ra + 1 < nr:
                    seen[r0 + ra + 1] += 1
            ci += GW
    assert (seen == 1).all()
