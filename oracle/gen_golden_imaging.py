"""Generate tests/golden/bicubic_cv2.npz: cv2.resize(..., INTER_CUBIC) outputs for the LR-synthesis shapes of the datasets
(build container only: needs opencv-python).      python oracle/gen_golden_imaging.py"""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "bicubic_cv2.npz")
CASES = [(160, 128, 40, 32), (40, 32, 160, 128), (161, 130, 40, 32), (100, 90, 33, 30), (52, 44, 208, 176), (48, 48, 24, 24),
         (9, 11, 27, 21)]
rng = np.random.default_rng(7)
out = {"cv2_version": np.array(cv2.__version__)}
for i, (hs, ws, hd, wd) in enumerate(CASES):
    img = rng.random((hs, ws), dtype=np.float32)
    out[f"src{i}"] = img
    out[f"dst{i}"] = cv2.resize(img, dsize=(wd, hd), interpolation=cv2.INTER_CUBIC)
np.savez_compressed(OUT, **out)
print("wrote", OUT, len(CASES), "cases, cv2", cv2.__version__)
