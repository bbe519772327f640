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
    for name in _CALLERS:
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "make_RDSTSR"):
            mod.make_RDSTSR = make_RDSTSR
            patched.append(name)
    return patched
