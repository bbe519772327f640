"""Dump the parameters the reference's own loader (utils/param_loader.py:6-21) reads from
config_files/RDST_E1_OASIS_example_SRx4.ini into tests/golden/e1_paras.json, so that bench.py and the tests can call
`make_RDSTSR(paras)` with exactly the reference's E1 configuration on machines where /root/reference is absent.

    python oracle/gen_paras.py        (needs /root/reference; test infrastructure, not product code)
"""
import json
import os
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
from utils.param_loader import ParametersLoader  # noqa: E402

INI = os.path.join(REF, "config_files", "RDST_E1_OASIS_example_SRx4.ini")
p = ParametersLoader(INI)
out = {}
for k in p.names:
    v = getattr(p, k)
    try:
        json.dumps(v)
    except TypeError:
        v = repr(v)
    out[k] = v
dst = os.path.join(ROOT, "tests", "golden", "e1_paras.json")
with open(dst, "w") as f:
    json.dump({"source": "config_files/RDST_E1_OASIS_example_SRx4.ini via utils/param_loader.ParametersLoader",
               "paras": out}, f, indent=1, sort_keys=True)
print("wrote", dst, len(out), "parameters")
