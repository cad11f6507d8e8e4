"""numpy model of the mixed-radix Stockham plans of gr-fosphor_b200/csrc/fft_power.cuh
(fft_power_kernel: pass structure, exchange indices, twiddle tables as engine.cu:build_twiddles
fills them) checked against numpy.fft for every plan, including the radix-64 two-pass ones.
Pins the index arithmetic of the kernels on the CPU; the arithmetic itself (f32, register DFTs)
is checked on the GPU by tests/test_gpu_engine_parity.py::test_fft_matches_double_dft."""
import numpy as np
import pytest

PLANS = [(512, 16, 32, 2), (1024, 32, 32, 2), (2048, 8, 16, 3), (4096, 16, 16, 3),
         (8192, 8, 32, 3), (16384, 16, 32, 3), (2048, 32, 64, 2), (4096, 64, 64, 2)]


def twiddles(n, r0, r1, npass):
    """engine.cu: build_twiddles"""
    t = np.arange(r1)[:, None]
    tw1 = np.exp(-2j * np.pi * t * np.arange(r0)[None, :] / (r0 * r1))            # [R1][R0]
    tw2 = None
    if npass == 3:
        p2 = r0 * r1
        tw2 = np.exp(-2j * np.pi * t * np.arange(p2)[None, :] / n)               # [R1][P2]
    return tw1, tw2


def model_fft(x, n, r0, r1, npass):
    nb0, nb1 = n // r0, n // r1
    tw1, tw2 = twiddles(n, r0, r1, npass)
    buf = np.zeros(n, complex)
    # pass 0 (P = 1, no twiddles): butterfly i takes x[i + t*NB0], writes buf[i*R0 + t] = X_R0[t]
    i = np.arange(nb0)
    v = x[i[:, None] + np.arange(r0)[None, :] * nb0]
    buf[(i[:, None] * r0 + np.arange(r0)[None, :]).ravel()] = np.fft.fft(v, axis=1).ravel()
    # pass 1 (P = R0)
    i = np.arange(nb1)
    k = i & (r0 - 1)
    v = buf[i[:, None] + np.arange(r1)[None, :] * nb1] * tw1[:, k].T
    y = np.fft.fft(v, axis=1)
    out = np.zeros(n, complex)
    if npass == 2:
        out[(i[:, None] + np.arange(r1)[None, :] * nb1).ravel()] = y.ravel()
        return out
    j = (i - k) * r1 + k
    buf2 = np.zeros(n, complex)
    buf2[(j[:, None] + np.arange(r1)[None, :] * r0).ravel()] = y.ravel()
    # pass 2 (P = R0 * R1), last
    p2 = r0 * r1
    k2 = i & (p2 - 1)
    v = buf2[i[:, None] + np.arange(r1)[None, :] * nb1] * tw2[:, k2].T
    out[(i[:, None] + np.arange(r1)[None, :] * nb1).ravel()] = np.fft.fft(v, axis=1).ravel()
    return out


@pytest.mark.parametrize("plan", PLANS)
def test_plan_matches_numpy_fft(plan):
    n = plan[0]
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    got = model_fft(x, *plan)
    ref = np.fft.fft(x)
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
