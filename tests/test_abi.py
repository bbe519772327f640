"""The C-ABI library must load and export every symbol declared in include/rdst_b200.h (no compute calls here)."""
import ctypes
import os
import re

import helpers


def _declared():
    src = open(os.path.join(helpers.ROOT, "include", "rdst_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rdst_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from rdst_b200 import _lib
    names = _declared()
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rdst_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == names          # python binding covers exactly the header


def test_library_reports_version_and_errors_without_gpu():
    from rdst_b200 import _lib
    lib = _lib.load()
    assert lib.rdst_abi_version() == _lib.ABI_VERSION
    # argument validation happens before any CUDA call, so it is testable on CPU
    rc = lib.rdst_window_attention_fwd(None, 0, None, None, 0, 1, 8, 8, 60, 6, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.rdst_last_error()
    rc = lib.rdst_conv3x3_fwd_bf16_tc(ctypes.c_void_p(16), 160, ctypes.c_void_p(16), ctypes.c_void_p(16), None, 0,
                                      ctypes.c_void_p(16), 64, 1, 8, 8, 96, 64, 1.0, 0, None)
    assert rc == -1 and b"supported (Cin,N)" in lib.rdst_last_error()


def test_no_oracle_import_in_product_package():
    pkg = os.path.join(helpers.ROOT, "rdst_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(root, f)).read()
                assert "rdst_oracle" not in txt and "abi_emulator" not in txt and "import oracle" not in txt, f
