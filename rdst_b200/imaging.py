"""LR synthesis and image metrics on the device -- the callers either side of the network.

The reference synthesises LR inputs and the bicubic "res" images with `cv2.resize(..., cv2.INTER_CUBIC)` on the host, one
slice at a time (datasets/basic_dataset.py:65-123, :258-301), and scores results with skimage PSNR / SSIM on the host
(metrics/sr_metrics.py:8-13).  Here whole volumes stay on the GPU: `resize_cubic` is cv2's INTER_CUBIC (same tap positions
and weights, within a few ulp of cv2's own SIMD arithmetic), `psnr` / `ssim` return one value per slice.
No CPU fallback: CUDA tensors only."""
import math

import numpy as np
import torch

from . import _lib

_TAPS = {}


def cubic_taps(n_src, n_dst):
    """Tap indices [n_dst][4] (int32, clamped = BORDER_REPLICATE) and weights [n_dst][4] (fp32) of cv2.INTER_CUBIC along one
    axis: source coordinate (d + 0.5) * n_src / n_dst - 0.5 evaluated in double, cubic convolution with A = -0.75, the
    fourth weight as 1 - the other three (OpenCV interpolateCubic)."""
    d = np.arange(n_dst, dtype=np.float64)
    scale = 1.0 / (float(n_dst) / float(n_src))
    f = (d + 0.5) * scale - 0.5
    s = np.floor(f)
    x = (f - s).astype(np.float32)
    A, one = np.float32(-0.75), np.float32(1)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    xm = one - x
    c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    idx = np.clip(s[:, None].astype(np.int64) + np.arange(-1, 3)[None, :], 0, n_src - 1).astype(np.int32)
    return idx, np.stack([c0, c1, c2, c3], axis=1).astype(np.float32)


def _taps_on(device, n_src, n_dst):
    key = (str(device), n_src, n_dst)
    if key not in _TAPS:
        idx, cf = cubic_taps(n_src, n_dst)
        _TAPS[key] = (torch.from_numpy(idx).to(device), torch.from_numpy(cf).to(device))
    return _TAPS[key]


def _images(x, what):
    if not x.is_cuda:
        raise RuntimeError(f"rdst_b200.imaging.{what}: input must be a CUDA tensor; this package has no CPU path")
    if x.dtype != torch.float32:
        raise TypeError(f"rdst_b200.imaging.{what}: float32 images expected, got {x.dtype}")
    if x.dim() == 4 and x.shape[1] != 1:
        raise ValueError(f"rdst_b200.imaging.{what}: single-channel images (B,1,H,W) expected, got {tuple(x.shape)}")
    if x.dim() not in (3, 4):
        raise ValueError(f"rdst_b200.imaging.{what}: expected (B,H,W) or (B,1,H,W), got {tuple(x.shape)}")
    return x.contiguous()


def resize_cubic(x, size):
    """cv2.resize(img, dsize=(size[1], size[0]), interpolation=cv2.INTER_CUBIC) for every image of a (B,1,H,W) / (B,H,W)
    CUDA tensor; `size` = (rows, cols) like MedicalImageBasicDataset.resize, or a float factor."""
    x = _images(x, "resize_cubic")
    hs, ws = x.shape[-2:]
    if isinstance(size, (int, float)):
        size = (size, size)
    hd, wd = size
    if isinstance(hd, float):
        hd, wd = int(hs * hd), int(ws * wd)          # basic_dataset.py:104-105
    if hd <= 0 or wd <= 0:
        raise ValueError("Size of output image should be positive")
    if (hd, wd) == (hs, ws):
        return x.clone()
    iy, cy = _taps_on(x.device, hs, hd)
    ix, cx = _taps_on(x.device, ws, wd)
    b = x.numel() // (hs * ws)
    out = torch.empty(x.shape[:-2] + (hd, wd), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("rdst_bicubic_resize_f32", _lib.ptr(x), _lib.ptr(out), _lib.ptr(ix), _lib.ptr(cx), _lib.ptr(iy), _lib.ptr(cy),
                  b, hs, ws, hd, wd, _lib.stream_ptr())
    return out


def make_test_pairs(hr, sr_scale):
    """Batched form of MIBasicValid.get_test_pair (datasets/basic_dataset.py:258-301) for one scale: from HR slices
    (B,1,H,W) returns (lr, gt, res): lr = cubic(HR -> H//s x W//s), gt = cubic(HR -> lr*s) (= HR when the sizes divide),
    res = cubic(lr -> gt size), the bicubic baseline / residual input."""
    hr = _images(hr, "make_test_pairs")
    h, w = hr.shape[-2:]
    lr = resize_cubic(hr, (int(h // sr_scale), int(w // sr_scale)))
    lh, lw = lr.shape[-2:]
    gt = resize_cubic(hr, (int(lh * sr_scale), int(lw * sr_scale)))
    res = resize_cubic(lr, tuple(gt.shape[-2:]))
    return lr, gt, res


def mse(gt, pred):
    """Per-image mean squared error in fp64 (B values on the device)."""
    gt, pred = _images(gt, "mse"), _images(pred, "mse")
    if gt.shape != pred.shape:
        raise ValueError("Input images must have the same dimensions.")
    b = gt.shape[0]
    n = gt.numel() // max(b, 1)
    out = torch.zeros(b, dtype=torch.float64, device=gt.device)
    with torch.cuda.device(gt.device):
        _lib.call("rdst_sqdiff_sum_f64", _lib.ptr(gt), _lib.ptr(pred), _lib.ptr(out), b, n, _lib.stream_ptr())
    return out / n


def psnr(gt, pred):
    """skimage peak_signal_noise_ratio(GT, P, data_range=1) per image (metrics/sr_metrics.py:8-9): 10 log10(1 / MSE)."""
    return 10.0 * torch.log10(1.0 / mse(gt, pred))


def ssim(gt, pred):
    """skimage structural_similarity(GT, P, data_range=1) per image (metrics/sr_metrics.py:12-13): 7x7 uniform window,
    sample covariance, mean over the valid region."""
    gt, pred = _images(gt, "ssim"), _images(pred, "ssim")
    if gt.shape != pred.shape:
        raise ValueError("Input images must have the same dimensions.")
    b = gt.shape[0]
    h, w = gt.shape[-2:]
    if min(h, w) < 7:
        raise ValueError("win_size exceeds image extent.")
    out = torch.zeros(b, dtype=torch.float64, device=gt.device)
    with torch.cuda.device(gt.device):
        _lib.call("rdst_ssim_sum_f64", _lib.ptr(gt), _lib.ptr(pred), _lib.ptr(out), b, h, w, _lib.stream_ptr())
    return out / float((h - 6) * (w - 6))
