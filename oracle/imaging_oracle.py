"""CPU restatement (numpy) of the imaging path either side of the network -- TEST INFRASTRUCTURE, never imported by the
product package.

* resize_cubic: cv2.resize(..., interpolation=cv2.INTER_CUBIC) as datasets/basic_dataset.py:65-123 calls it.  The
  algorithm lives in OpenCV (opencv-python 4.13.0 in the build container), not in /root/reference: restated from its
  published resize: source coordinate (d+0.5)*scale-0.5 in double, interpolateCubic weights with A = -0.75, replicated
  borders, horizontal pass then vertical pass in fp32.  PINNED: tests/golden/bicubic_*.npz are cv2 outputs generated here
  (oracle/gen_golden_imaging.py); this restatement agrees with them to <= 7e-7 (cv2's own SIMD/FMA order is build-dependent
  at the 1-ulp level, so bit-exactness is not defined).
* psnr / ssim: skimage.metrics.peak_signal_noise_ratio / structural_similarity(data_range=1) as metrics/sr_metrics.py:8-13
  calls them.  scikit-image is NOT installed in the build container and the reference ships no metric fixtures: PSNR is the
  closed form 10 log10(1/MSE); SSIM restates skimage's published algorithm (uniform 7x7 filter via scipy.ndimage, sample
  covariance, K1 = 0.01, K2 = 0.03, crop of 3) -- PARITY UNPINNED for SSIM."""
import numpy as np

f32 = np.float32


def cubic_taps(n_src, n_dst):
    d = np.arange(n_dst, dtype=np.float64)
    scale = 1.0 / (float(n_dst) / float(n_src))
    f = (d + 0.5) * scale - 0.5
    s = np.floor(f)
    x = (f - s).astype(f32)
    A, one = f32(-0.75), f32(1)
    c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
    xm = one - x
    c2 = ((A + f32(2)) * xm - (A + f32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    idx = np.clip(s[:, None].astype(np.int64) + np.arange(-1, 3)[None, :], 0, n_src - 1)
    return idx, np.stack([c0, c1, c2, c3], axis=1).astype(f32)


def resize_cubic(img, hd, wd):
    """img: [H][W] float32 -> [hd][wd] float32."""
    img = np.asarray(img, dtype=f32)
    hs, ws = img.shape
    iy, cy = cubic_taps(hs, hd)
    ix, cx = cubic_taps(ws, wd)
    g = img[:, ix]                                                    # [hs][wd][4]
    h = g[:, :, 0] * cx[None, :, 0]
    for k in range(1, 4):
        h = h + g[:, :, k] * cx[None, :, k]
    v = h[iy]                                                         # [hd][4][wd]
    out = v[:, 0] * cy[:, 0, None]
    for k in range(1, 4):
        out = out + v[:, k] * cy[:, k, None]
    return out.astype(f32)


def psnr(gt, p):
    mse = np.mean((np.asarray(gt, np.float64) - np.asarray(p, np.float64)) ** 2)
    return 10.0 * np.log10(1.0 / mse)


def ssim(gt, p):
    from scipy.ndimage import uniform_filter
    x, y = np.asarray(gt, np.float64), np.asarray(p, np.float64)
    win, NP = 7, 49.0
    cov_norm = NP / (NP - 1.0)
    ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
    uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean())
