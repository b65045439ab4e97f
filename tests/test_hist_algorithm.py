"""The ALGORITHM of csrc/hist.cu restated in numpy (telescoped cdf over the 257 bin edges, +-6.8-bin cut-off,
32.32 fixed-point accumulation) against the oracle's literal restatement of losses.py:8-87.  CPU only: it pins the
mathematics the kernel relies on (the kernel itself is checked on the GPU in tests/test_gpu_losses.py)."""
import numpy as np
import pytest
import torch

from oracle import restate as R

K, CUT = 256, 6.8


def edge_sums_fixed_point(x):
    """G_j = sum_i sigmoid(2.5 (256 x_i - j)), j = 0..256, as 32.32 fixed point, the way hist_acc_kernel forms it."""
    t = np.clip(x.astype(np.float32) * np.float32(K), -16.0, K + 16.0).astype(np.float64)
    acc = np.zeros(K + 1, dtype=object)
    jlo = np.floor(t - CUT).astype(np.int64)
    jlo = np.minimum(jlo, K)
    cnt = np.zeros(K + 1, dtype=np.int64)
    np.add.at(cnt, jlo[jlo >= 0], 1)
    hard = np.cumsum(cnt[::-1])[::-1]                      # elements with jlo >= j count 1.0 for edge j
    jhi = np.minimum(np.ceil(t + CUT).astype(np.int64), K)
    for ti, lo, hi in zip(t, jlo, jhi):
        for j in range(max(lo + 1, 0), hi + 1):
            s = 1.0 / (1.0 + np.exp(-2.5 * (ti - j)))
            acc[j] += int(round(s * 4294967296.0))
    return [int(a) + (int(h) << 32) for a, h in zip(acc, hard)]


def emd_loss(x, y):
    B = x.shape[0]
    norm_x, norm_y = x.shape[1] * x.shape[2], y.shape[1] * y.shape[2]      # losses.py:54
    total = 0.0
    for b in range(B):
        gx, gy = edge_sums_fixed_point(x[b].ravel()), edge_sums_fixed_point(y[b].ravel())
        cx = np.array([(gx[0] - gx[t + 1]) / 4294967296.0 / norm_x for t in range(K)])
        cy = np.array([(gy[0] - gy[t + 1]) / 4294967296.0 / norm_y for t in range(K)])
        total += float(((cx - cy) ** 2).sum())
    return total / B


@pytest.mark.parametrize("shape,seed", [((2, 3, 6, 5), 1), ((1, 3, 9, 4), 2)])
def test_telescoped_fixed_point_emd_matches_the_reference_formulation(shape, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(*shape, generator=g) * 1.3 - 0.15            # a few values outside [0, 1]
    y = torch.rand(*shape, generator=g) ** 2
    ref = R.compute_hist_loss(x.double(), y.double()).item()
    got = emd_loss(x.numpy(), y.numpy())
    assert got == pytest.approx(ref, rel=2e-5)


def test_identical_inputs_give_exactly_zero_and_order_does_not_matter():
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 5, 5, generator=g).numpy()
    assert emd_loss(x, x) == 0.0
    flat = x[0].ravel()
    perm = np.random.default_rng(0).permutation(flat.size)
    assert edge_sums_fixed_point(flat) == edge_sums_fixed_point(flat[perm])        # integer accumulation: bit-deterministic
