"""LR synthesis (cv2 INTER_CUBIC) and PSNR / SSIM -- SURVEY 8f row 4.  CPU: the numpy oracle against cv2 golden outputs
(generated in the build container) and its own properties; GPU: the kernels against the oracle (bit-exact for the resize:
same tap tables, same fp32 operation order) and against the cv2 goldens."""
import os

import numpy as np
import pytest
import torch

import helpers
import imaging_oracle as IO

G = np.load(os.path.join(helpers.GOLDEN, "bicubic_cv2.npz"))
NCASE = len([k for k in G.files if k.startswith("src")])
# cv2's own SIMD/FMA summation order is build-dependent at the 1-ulp level: parity with cv2 is a tolerance, not bit-exactness
CV2_TOL = 1e-6


@pytest.mark.parametrize("i", range(NCASE))
def test_oracle_matches_cv2_golden(i):
    src, dst = G[f"src{i}"], G[f"dst{i}"]
    y = IO.resize_cubic(src, *dst.shape)
    assert y.shape == dst.shape and y.dtype == np.float32
    assert np.abs(y - dst).max() < CV2_TOL


def test_tap_tables_match_oracle_and_properties():
    from rdst_b200 import imaging
    for n_src, n_dst in ((160, 40), (40, 160), (161, 40), (9, 27)):
        i0, c0 = imaging.cubic_taps(n_src, n_dst)
        i1, c1 = IO.cubic_taps(n_src, n_dst)
        assert (i0 == i1).all() and (c0 == c1).all() and i0.dtype == np.int32
        assert np.abs(c0.sum(1) - 1).max() < 2e-7 and i0.min() >= 0 and i0.max() <= n_src - 1
    i, c = imaging.cubic_taps(160, 40)                   # x4 down-sampling: taps at 4d .. 4d+3, weights of the half-pixel phase
    assert (i[:, 0] == 4 * np.arange(40)).all() and np.allclose(c, [[-0.09375, 0.59375, 0.59375, -0.09375]])


def test_metric_oracle_properties():
    rng = np.random.default_rng(0)
    a = rng.random((40, 48)).astype(np.float32)
    assert IO.ssim(a, a) == pytest.approx(1.0, abs=1e-12)
    b = np.clip(a + 0.05 * rng.standard_normal(a.shape).astype(np.float32), 0, 1)
    assert 0 < IO.ssim(a, b) < 1 and IO.ssim(a, b) == pytest.approx(IO.ssim(b, a), abs=1e-12)
    assert IO.psnr(a, a + np.float32(0.1)) == pytest.approx(20.0, abs=1e-4)


def test_cpu_inputs_are_rejected():
    from rdst_b200 import imaging
    with pytest.raises(RuntimeError, match="no CPU path"):
        imaging.resize_cubic(torch.zeros(1, 1, 8, 8), (4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        imaging.psnr(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(NCASE))
def test_gpu_resize_matches_oracle_and_cv2(i):
    from rdst_b200 import imaging
    src, dst = G[f"src{i}"], G[f"dst{i}"]
    x = torch.from_numpy(np.stack([src, src[::-1].copy(), src * 0.5])).cuda()           # a batch of three images
    y = imaging.resize_cubic(x[:, None], dst.shape).cpu().numpy()[:, 0]
    for k, s in enumerate((src, src[::-1].copy(), src * np.float32(0.5))):
        assert (y[k] == IO.resize_cubic(s, *dst.shape)).all()                           # bit-exact vs the oracle
    assert np.abs(y[0] - dst).max() < CV2_TOL                                           # and within cv2's own ulp noise


@pytest.mark.gpu
def test_gpu_test_pairs_and_metrics():
    from rdst_b200 import imaging
    rng = np.random.default_rng(3)
    hr = rng.random((5, 1, 160, 128)).astype(np.float32)
    lr, gt, res = imaging.make_test_pairs(torch.from_numpy(hr).cuda(), 4.0)
    assert lr.shape == (5, 1, 40, 32) and gt.shape == (5, 1, 160, 128) and res.shape == (5, 1, 160, 128)
    assert torch.equal(gt.cpu(), torch.from_numpy(hr))                                  # same size: no resampling (:116-117)
    for b in range(5):
        assert (lr[b, 0].cpu().numpy() == IO.resize_cubic(hr[b, 0], 40, 32)).all()
    p, s = imaging.psnr(gt, res).cpu().numpy(), imaging.ssim(gt, res).cpu().numpy()
    resn = res.cpu().numpy()
    for b in range(5):
        assert p[b] == pytest.approx(IO.psnr(hr[b, 0], resn[b, 0]), abs=1e-9)
        assert s[b] == pytest.approx(IO.ssim(hr[b, 0], resn[b, 0]), abs=1e-9)
    # odd sizes (valid region not a multiple of the block shape) and the full-size property PSNR(x, x) = inf
    a = torch.from_numpy(rng.random((2, 37, 53)).astype(np.float32)).cuda()
    b2 = (a + 0.01).clamp(0, 1)
    s2 = imaging.ssim(a, b2).cpu().numpy()
    for k in range(2):
        assert s2[k] == pytest.approx(IO.ssim(a[k].cpu().numpy(), b2[k].cpu().numpy()), abs=1e-9)
    assert torch.isinf(imaging.psnr(a, a)).all()
