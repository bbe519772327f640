"""Make the reference code base use rdst_b200 without editing it.

The reference trainer/tester do `from networks.swinIR_variations import make_RDSTSR`
(models/trans_sr_trainer.py:3, models/trans_sr_tester.py:3).  `install()` rebinds that name (and the copy in
networks.rdst_variations) to this package's factory, and patches trainer/tester modules that were already
imported.  Call it once before constructing TransSRTrainer / TransSRTester, e.g. from a sitecustomize or at the
top of train.py / test.py:   `import rdst_b200; rdst_b200.install()`.
"""
import importlib
import sys

from .network import RDSTSR, make_RDSTSR
from .swinir import SwinIR, swinir_make_model

_TARGETS = ("networks.swinIR_variations", "networks.rdst_variations")
_CALLERS = ("models.trans_sr_trainer", "models.trans_sr_tester")


def install(strict=False):
    """Returns the list of module names that were patched."""
    patched = []
    for name in _TARGETS:
        mod = sys.modules.get(name)
        if mod is None:
            try:
                mod = importlib.import_module(name)
            except Exception:
                if strict:
                    raise
                continue
        mod.make_RDSTSR = make_RDSTSR
        mod.RDSTSR = RDSTSR
        patched.append(name)
    # vanilla SwinIR (feature_generator = 'swinir'): `from networks.swin_transformer_sr import swinir_make_model`
    # (models/trans_sr_trainer.py:2, models/trans_sr_tester.py:2)
    mod = sys.modules.get("networks.swin_transformer_sr")
    if mod is None:
        try:
            mod = importlib.import_module("networks.swin_transformer_sr")
        except Exception:
            if strict:
                raise
            mod = None
    if mod is not None:
        mod.swinir_make_model = swinir_make_model
        mod.SwinIR = SwinIR
        patched.append("networks.swin_transformer_sr")
    for name in _CALLERS:
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "make_RDSTSR"):
            mod.make_RDSTSR = make_RDSTSR
            if hasattr(mod, "swinir_make_model"):
                mod.swinir_make_model = swinir_make_model
            patched.append(name)
    return patched
