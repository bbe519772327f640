"""bench.py's reference arm runs on CPU (the oracle on the host cores): check the JSON contract of its line here; the GPU arm
prints the same keys plus roofline / clocks / gpu_launches and is exercised on the B200 by the driver."""
import json
import os
import subprocess
import sys

import helpers


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, RDST_BENCH_REF_BUDGET_S="15"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("HR output Mpix/s") and d["scaling"] == "weak" and d["vs_baseline"] is None
    # "reference" = the real networks/rdst_variations.py (importable in the build container), "port" = the oracle (GPU box)
    assert d["cpu_baseline"]["kind"] == ("reference" if os.path.isdir("/root/reference/networks") else "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
