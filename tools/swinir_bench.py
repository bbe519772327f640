"""SwinIR-lite x4 (the [SwinIR] section of RDST_E1_OASIS_example_SRx4.ini: 4 RSTBs x 6 Swin blocks, C = 60) on the RDST kernels:
HR Mpix/s at the cfg2 batch (176 LR slices 40x32 per step), bf16, inputs resident, L2 flushed between steps, CUDA events;
the oracle (PyTorch-CPU restatement of the reference SwinIR) on 8 slices for scale."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
import swinir_oracle as SO  # noqa: E402

c = helpers.load_swinir_case("swinir_ini_x4_40x32")
m = helpers.make_swinir(c, "bf16").cuda().eval()
m.load_state_dict(c["sd"])
x = torch.rand(176, 1, 40, 32, generator=torch.Generator().manual_seed(1)).cuda()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = m(x)
    tot, n = 0.0, 10
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        y = m(x)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
ms = tot / n
ref = SO.forward(c["sd"], x[:8].cpu(), 4)
err = (y[:8].cpu() - ref).abs().max().item()
torch.set_num_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
SO.forward(c["sd"], x[:8].cpu(), 4)
cpu_s = time.perf_counter() - t0
flop_per_lr_px = 24 * 72960 + 4 * 64800 + 1080 + 64800 + 2 * 9 * 60 * 16        # Swin blocks (16C^2+256C), RSTB convs, head, cab, upsample
print(json.dumps({"what": "SwinIR-lite x4 inference on librdst_b200 (bf16), 176 LR slices 40x32 per step",
                  "ms_per_step": round(ms, 3), "hr_mpix_per_s": round(176 * 160 * 128 / ms / 1e3, 1),
                  "tflops_algorithmic": round(flop_per_lr_px * 176 * 40 * 32 / ms / 1e9, 1),
                  "max_abs_err_vs_oracle_fp32": err, "oracle_output_max_abs": ref.abs().max().item(),
                  "cpu_oracle_hr_mpix_per_s": round(8 * 160 * 128 / cpu_s / 1e6, 3), "cpu_threads": os.cpu_count()}))
