"""Boundary tests against the REAL reference code base (build container only: skipped where /root/reference is absent).

rdst_b200.install() must rebind the factories the reference trainer / tester import (models/trans_sr_trainer.py:2-3,
models/trans_sr_tester.py:2-3), the module built from the reference's own ini through the reference's own loader must have
the reference module's state_dict (keys, order, shapes, dtypes; strict loads both ways), and the tester's inference loop
(models/basic_tester.py:104-115 + models/trans_sr_tester.py:124-166) must give the reference's results with the drop-in
module behind it -- kernels replaced by their contract restatements (tests/abi_emulator.py), the reference module on CPU
as the oracle."""
import importlib
import importlib.machinery
import os
import sys
import types

import numpy as np
import pytest
import torch

import helpers
from abi_emulator import emulated_abi

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "networks")), reason="reference tree not present")


class _Anything(types.ModuleType):
    """Stand-in for a third-party module the reference imports at module level but this path never calls."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        raise RuntimeError(f"stubbed dependency {self.__name__} was called")


@pytest.fixture(scope="module")
def ref_modules():
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for p in (os.path.join(helpers.ROOT, "oracle", "_shim"), REF):
        sys.path.insert(0, p)
    # an unrelated installed package named `datasets` / `metrics` / `models` may already be imported: the reference's
    # top-level packages must win inside this fixture (everything is restored afterwards)
    for k in list(sys.modules):
        if k.split(".")[0] in ("networks", "models", "utils", "datasets", "metrics", "loss"):
            del sys.modules[k]
    for k in ("networks", "models", "utils", "datasets", "metrics", "loss"):        # the reference's packages have no
        pkg = types.ModuleType(k)                                                   # __init__.py: an installed regular
        pkg.__path__ = [os.path.join(REF, k)]                                       # package of the same name (HF
        pkg.__spec__ = importlib.machinery.ModuleSpec(k, None, is_package=True)     # `datasets`) would shadow them
        sys.modules[k] = pkg
    stubbed = []

    def stub(name):
        m = _Anything(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        m.__path__ = []
        sys.modules[name] = m
        stubbed.append(name)

    def imp(name):
        # third-party modules the reference imports at module level (skimage, sewar, nibabel, ...) are absent here and are
        # never called on this path: stub whatever is missing, one ModuleNotFoundError at a time
        for _ in range(64):
            try:
                return importlib.import_module(name)
            except ModuleNotFoundError as e:
                if e.name is None or e.name.split(".")[0] in ("networks", "models", "utils", "datasets", "metrics", "loss"):
                    raise
                stub(e.name)
        raise RuntimeError("too many missing modules")

    rv = imp("networks.rdst_variations")
    ref_make = rv.make_RDSTSR                       # the reference factory, before install() rebinds the name
    sv = imp("networks.swinIR_variations")
    st = imp("networks.swin_transformer_sr")
    ref_swinir_make = st.swinir_make_model
    tester_mod = imp("models.trans_sr_tester")
    imp("models.trans_sr_trainer")
    del sys.modules["models.trans_sr_trainer"]      # re-imported after install() by the first test
    from utils.param_loader import ParametersLoader
    paras = ParametersLoader(os.path.join(REF, "config_files", "RDST_E1_OASIS_example_SRx4.ini"))
    ref_cls = rv.RDSTSR
    torch.manual_seed(0)
    ref_model = ref_make(paras)                     # built BEFORE install(): afterwards the names resolve to rdst_b200
    assert type(ref_model) is ref_cls and type(ref_model).__module__ == "networks.rdst_variations"
    import rdst_b200
    patched = rdst_b200.install(strict=True)
    yield types.SimpleNamespace(rv=rv, sv=sv, st=st, tester_mod=tester_mod, paras=paras, ref_model=ref_model, ref_cls=ref_cls,
                                ref_swinir_make=ref_swinir_make, patched=patched)
    # undo: other tests must see pristine modules
    for k in list(sys.modules):
        if k not in saved_mods and (k.split(".")[0] in ("networks", "models", "utils", "datasets", "metrics", "loss", "timm") or
                                     k in stubbed):
            del sys.modules[k]
    sys.path[:] = saved_path
    for k, v in saved_mods.items():
        sys.modules.setdefault(k, v)


def test_install_rebinds_reference_factories(ref_modules):
    import rdst_b200
    r = ref_modules
    assert r.sv.make_RDSTSR is rdst_b200.make_RDSTSR and r.rv.make_RDSTSR is rdst_b200.make_RDSTSR
    assert r.tester_mod.make_RDSTSR is rdst_b200.make_RDSTSR               # imported before install(): patched in place
    assert r.st.swinir_make_model is rdst_b200.swinir_make_model and r.tester_mod.swinir_make_model is rdst_b200.swinir_make_model
    assert "models.basic_tester.BasicTester.inference" in r.patched
    import models.trans_sr_trainer as trainer_mod                          # imported after install(): sees the new names
    assert trainer_mod.make_RDSTSR is rdst_b200.make_RDSTSR


def test_factory_on_reference_ini_has_reference_state_dict(ref_modules):
    import rdst_b200
    r = ref_modules
    torch.manual_seed(0)
    ours = r.sv.make_RDSTSR(r.paras)                 # what TransSRTrainer / TransSRTester call after install()
    assert isinstance(ours, rdst_b200.RDSTSR) and ours.sr_scale == 4
    import copy
    ref = copy.deepcopy(r.ref_model)
    assert not isinstance(ref, rdst_b200.RDSTSR) and isinstance(ref, r.ref_cls)
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o.keys()) == list(sd_r.keys()) and len(sd_o) == 826
    for k in sd_r:
        assert sd_o[k].shape == sd_r[k].shape and sd_o[k].dtype == sd_r[k].dtype, k
    ours.load_state_dict(sd_r, strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    assert [p.requires_grad for p in ours.parameters()] == [p.requires_grad for p in ref.parameters()]
    # the committed paras fixture (used by bench.py where the reference tree is absent) is what the loader reads
    import json
    with open(os.path.join(helpers.ROOT, "tests", "golden", "e1_paras.json")) as f:
        fixture = json.load(f)["paras"]
    for k in r.paras.names:
        v = getattr(r.paras, k)
        try:
            v = json.loads(json.dumps(v))
        except TypeError:
            v = repr(v)
        assert fixture[k] == v, k


class _DS:
    """Minimal stand-in for datasets.OASIS_dataset.OASISMultiSRTest: test pairs in the documented format (:276-293)."""

    def __init__(self, slices):
        self.slices = slices

    def test_len(self):
        return len(self.slices)

    def get_test_pair(self, i):
        lr = self.slices[i]
        return {4.0: {"in": lr, "res": lr, "gt": np.zeros((lr.shape[2] * 4, lr.shape[3] * 4, 1), np.float32), "sr_factor": 4.0,
                      "real_sr_scale": 4.0}}


def _tester(r, model):
    t = object.__new__(r.tester_mod.TransSRTester)       # no datasets / metrics / output folders: only the inference path
    t.device = torch.device("cpu")
    t.sr_generator = "rdst"
    t.single_scale_model = model
    t.batch_size = 2
    t.model_input_with_scale_flag = "no"
    return t


def test_tester_inference_matches_reference_module(ref_modules):
    """TransSRTester with the drop-in module behind it (batched loop installed) == the reference tester's own per-slice
    loop over the reference module, same weights, same slices."""
    r = ref_modules
    import copy
    import rdst_b200
    ref = copy.deepcopy(r.ref_model).eval()
    assert not isinstance(ref, rdst_b200.RDSTSR)
    from synth_weights import fill_state_dict
    sd = fill_state_dict(ref.state_dict(), 7, True)
    ref.load_state_dict(sd, strict=True)
    ours = r.sv.make_RDSTSR(r.paras)
    ours.load_state_dict(sd, strict=True)
    ours.forward = lambda x, sr_scale=None, out=None: ours._exec._forward_impl(x)      # CPU: kernels = contract restatements
    g = torch.Generator().manual_seed(3)
    slices = [torch.rand(1, 1, 16, 24, generator=g) for _ in range(3)] + [torch.rand(1, 1, 8, 16, generator=g)]
    ds = _DS(slices)
    import models.basic_tester as bt
    assert getattr(bt.BasicTester.inference, "_rdst_b200_batched", False)
    with emulated_abi():
        preds, samples = _tester(r, ours).inference(ds, return_sample=True)
    # reference behaviour: the original loop (one __inference_one__ per pair) over the reference module
    t_ref = _tester(r, ref)
    expect = [t_ref.__inference_one__(ds.get_test_pair(i)) for i in range(ds.test_len())]
    assert len(preds) == len(expect) == 4 and len(samples) == 4
    for p, e in zip(preds, expect):
        assert list(p.keys()) == list(e.keys()) == [4.0]
        assert p[4.0].shape == e[4.0].shape and p[4.0].dtype == e[4.0].dtype
        assert np.abs(p[4.0] - e[4.0]).max() < 2e-5
    # the reference per-slice entry point still works with the drop-in module (it is what TransSRTrainer validation calls)
    with emulated_abi():
        one = _tester(r, ours).__inference_one__(ds.get_test_pair(0))
    assert np.abs(one[4.0] - expect[0][4.0]).max() < 2e-5
