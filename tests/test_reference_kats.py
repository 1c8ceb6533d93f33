"""Inline known answers of the reference's own unit tests, evaluated on the mirrors of this repository (CPU):
tests/test_clip.py (clipping of local energies and wave-function ratios) and tests/test_overlap.py (symmetrisation and
weighted mean of the overlap matrix) of deepqmc/deepqmc.  Same inputs, same expected numbers."""
import math

import torch

from deepqmc_b200.energy import clip_local_energy, median_clip_and_mask, median_log_squeeze_and_mask
from deepqmc_b200.overlap import (clip_psi_ratio, compute_mean_overlap, psi_ratio_clip_and_mask,
                                  symmetrize_overlap_with_clipped_geometric_mean)

T = lambda x: torch.tensor(x, dtype=torch.float64)


def _trivial(x):
    return torch.clamp(x, -1.0, 1.0), torch.ones_like(x, dtype=torch.bool)


def test_vmapped_clipping():  # tests/test_clip.py TestVmappedClipping
    c, m = clip_local_energy(_trivial, T([[[0.5, 2.0, -3.0]]]))
    assert torch.allclose(c, T([[[0.5, 1.0, -1.0]]])) and m.all()
    c, m = clip_psi_ratio(_trivial, T([[[[0.5, 2.0, -3.0]]]]))
    assert torch.allclose(c, T([[[[0.5, 1.0, -1.0]]]])) and m.all()


def test_median_clip_outlier_is_clipped_and_masked():  # TestMedianClipAndMask
    c, m = median_clip_and_mask(T([1.0, 2.0, 3.0, 4.0, 100.0]), clip_width=1.0, median_center=True, exclude_width=5.0)
    assert torch.allclose(c, T([1.0, 2.0, 3.0, 4.0, 23.2])) and m.tolist() == [True, True, True, True, False]


def test_median_log_squeeze():  # TestMedianLogSqueezeAndMask
    x = T([1.0, 2.0, 3.0, 4.0, 100.0])
    s, m = median_log_squeeze_and_mask(x, clip_width=1.0, quantile=0.95)
    assert all(abs(s[i].item() - x[i].item()) < 1e-8 * 1e5 and torch.allclose(s[i], x[i], rtol=1e-5, atol=1e-8 * 1e3) for i in (1, 2, 3))
    width = 78.0
    z = 97.0 / (2 * width)
    expected = 3.0 + 2 * width * math.copysign(math.log1p((abs(z) + 0.5 * z**2 + abs(z) ** 3) / (1 + z**2)), z)
    assert abs(s[4].item() - expected) < 1e-8 * expected and 3.0 + width < s[4].item() < 3.0 + 2 * width and m.all()
    _, m = median_log_squeeze_and_mask(x, clip_width=1.0, quantile=0.95, exclude_width=1.0)
    assert m.tolist() == [True, True, True, True, False]
    s, m = median_log_squeeze_and_mask(T([1.0, 2.0, 3.0, 4.0, 5.0]))
    assert s.shape == (5,) and m.all()


def test_psi_ratio_clip_degenerate_sigma():  # TestPsiRatioClipAndMask
    c, m = psi_ratio_clip_and_mask(T([1.0, 1.0, 1.0, 1.0, 10.0]), clip_width=2.0, exclude_width=3.0)
    assert torch.allclose(c, torch.ones(5, dtype=torch.float64)) and m.tolist() == [True, True, True, True, False]


def test_symmetrize_overlap():  # tests/test_overlap.py TestSymmetrizeOverlap
    f = symmetrize_overlap_with_clipped_geometric_mean
    s06, s6, s02 = math.sqrt(0.06), math.sqrt(6.0), math.sqrt(0.2)
    assert torch.allclose(f(T([[1.0, 0.3], [0.2, 1.0]])), T([[1.0, s06], [s06, 1.0]]))
    y = f(T([[1.0, -0.4], [0.3, 1.0]]))
    assert y[0, 1].abs().item() == 0.0 and y[1, 0].item() == 0.0 and y[0, 0].item() == 1.0 and y[1, 1].item() == 1.0
    assert torch.allclose(f(T([[1.0, 2.0], [3.0, 1.0]])), T([[1.0, s6], [s6, 1.0]]))
    x = T([[1.0, 0.3, -0.5], [0.2, 1.0, 0.4], [0.6, 0.5, 1.0]])
    assert torch.allclose(f(x), T([[1.0, s06, 0.0], [s06, 1.0, s02], [0.0, s02, 1.0]]))


def test_mean_overlap_weighted_mean_and_symmetrization():  # TestComputeMeanOverlap (one molecule of its batch)
    psi_ratio = T([[[1.0, 1.0], [0.2, 0.4]], [[0.3, 0.5], [1.0, 1.0]]])
    weight = T([[1.0, 1.0], [0.8, 1.2]])
    loss, stats = compute_mean_overlap(psi_ratio, weight)
    s = math.sqrt(0.128)
    assert abs(loss.item() - 0.128) < 1e-12 and torch.allclose(stats['overlap/pairwise/mean'], T([[1.0, s], [s, 1.0]]))
