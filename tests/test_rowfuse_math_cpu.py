"""The arithmetic of the attention fusions (conv_gemm.cuh, FLAVOR_ROWFUSE), restated in numpy as the
epilogue performs it — per N tile an online (max, sum exp) over 32-column chunks with the maximum kept in the log2 domain
(exponentials are ex2(v * log2e - m * log2e)), then the combination of the tiles' partials — against a plain softmax; and
the flash-attention identity the fused dS relies on, rowsum(dP o P) = dO . O. Checks the formulas, not the kernel (that is
tests/test_kernel_options_gpu.py)."""
import numpy as np


def test_two_pass_softmax_formulas():
    rng = np.random.RandomState(0)
    rows, N, BN, CH = 7, 1024, 128, 32
    S = (rng.randn(rows, N) * 6).astype(np.float32)
    nt = N // BN
    stat = np.zeros((rows, nt, 2), np.float32)
    log2e = np.float32(1.4426950408889634)
    for t in range(nt):                      # pass 1, one tile at a time, chunks of CH columns
        rmax = np.full(rows, -np.inf, np.float32)   # log2e * running maximum
        rsum = np.zeros(rows, np.float32)
        for c in range(0, BN, CH):
            v = S[:, t * BN + c: t * BN + c + CH]
            nm = np.maximum(rmax, v.max(1) * log2e)
            rsum = rsum * np.exp2(rmax - nm) + np.exp2(v * log2e - nm[:, None]).sum(1)
            rmax = nm
        stat[:, t, 0], stat[:, t, 1] = rmax, rsum
    M = stat[:, :, 0].max(1)                 # pass 2 prologue (log2 domain)
    L = (stat[:, :, 1] * np.exp2(stat[:, :, 0] - M[:, None])).sum(1)
    P = np.exp2(S * log2e - M[:, None]) / L[:, None]
    ref = np.exp(S - S.max(1, keepdims=True))
    ref /= ref.sum(1, keepdims=True)
    assert np.abs(P - ref).max() < 2e-6 and np.abs(P.sum(1) - 1).max() < 1e-5


def test_fused_ds_identity():
    rng = np.random.RandomState(1)
    Nq, Nk, dv = 12, 40, 16
    P = rng.rand(Nq, Nk); P /= P.sum(1, keepdims=True)
    g = rng.randn(Nk, dv)
    dO = rng.randn(Nq, dv)
    O = P @ g
    dP = dO @ g.T
    D = (dO * O).sum(1)
    np.testing.assert_allclose((dP * P).sum(1), D, rtol=1e-10)
    dS_fused = P * (dP - D[:, None])
    dS_ref = P * (dP - (dP * P).sum(1, keepdims=True))   # softmax_bwd_kernel
    np.testing.assert_allclose(dS_fused, dS_ref, rtol=1e-9, atol=1e-12)
