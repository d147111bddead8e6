"""CPU emulation of the index arithmetic of the staged expand kernel in csrc/kernels_hbm.cuh (the kernel itself is
checked on the GPU by the parity suite): the gather index ((b*T + t*stride + kk)*J + j)*Fin + i that fills the staged
rows is the input window of the strided (k,1) convolution it implements (gast_net.py:163-164)."""
import numpy as np
import pytest


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
