"""rdst_b200 -- Blackwell-native (sm_100a) implementation of the RDST super-resolution network hot path.

Public surface (mirrors the reference's networks/rdst_variations.py):
    RDSTSR, make_RDSTSR      drop-in nn.Module / factory (same ctor args, forward(x), state_dict keys)
    RDSTSR_N, ESTSR          the global-bottleneck and the residual-in-residual variants of the same file
    SwinIR, swinir_make_model  the reference's vanilla SwinIR (lightweight configuration) on the same kernels
    install()                rebinds the reference's `networks.*` factories to this implementation
    imaging                  LR synthesis (cv2 INTER_CUBIC) and PSNR / SSIM on the device (datasets / metrics either side)
All arithmetic runs in librdst_b200.so (include/rdst_b200.h); importing this package without the built
library works (so that state_dicts can be inspected), but the first forward raises.
"""
from .network import ESTSR, RDSTSR, RDSTSR_N, make_RDSTSR  # noqa: F401
from .swinir import SwinIR, swinir_make_model  # noqa: F401
from .install import install  # noqa: F401
from . import imaging  # noqa: F401

__all__ = ["RDSTSR", "RDSTSR_N", "ESTSR", "make_RDSTSR", "SwinIR", "swinir_make_model", "install", "imaging"]
__version__ = "0.1.0"
