"""bench.py -- RDST-E1 x4 super-resolution throughput on B200 (HR output megapixels / second).

Workload (BASELINE.json configs[1] / SURVEY 8d cfg2): RDST-E1 x4, bf16 mode, one synthetic OASIS-shaped volume
per GPU per step = 176 LR slices of 1x40x32 -> 176 x 160x128 HR pixels (3.60 Mpix), random-init weights.
Multi-GPU: every rank super-resolves its own volume (slices are independent; no data-path collective) => weak scaling.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]
                    [--no-extras] [--no-cpu-baseline]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` times the forward with inputs resident in HBM; `e2e` times the module's
public call (`rdst_b200.make_RDSTSR(paras)` with the reference's E1 paras) with pinned HOST input and a device->host
read of the HR result inside the timed region.  Before anything is timed, sampled slices of the bench batch are
checked against the CPU oracle (`parity_max_abs`; the run aborts above the north_star tolerance).  Extra keys (not the
headline): `rdst_e_cfg3` = the 4-RDSTB "RDST-E" variant on the same batch; `rdst_e_cfg3_strong` = ONE such volume split
over the N ranks, one CUDA-graph replay per rank (strong scaling); `train_cfg4` = one data-parallel training
step (32 x 1x24x24 per GPU, L1, Adam, NCCL gradient all-reduce inside the step's CUDA graph when N > 1);
`roofline_hbm_kernel` = the step's HBM-bound kernel (reconstruction conv) against the measured HBM peak, next to
`roofline` (the dominant kernel, fused window attention at C = 120, against the measured bf16 tensor peak).
`--impl reference` times the reference's own module on the host cores (the real networks/rdst_variations.py when
/root/reference is importable, else the oracle port) on the full 176-slice batch.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

import types  # noqa: E402

SLICES, LR_H, LR_W, SCALE = 176, 40, 32, 4
PARITY_SLICES = [0, 59, 117, 175]                 # sampled slices checked against the oracle before timing


def e1_paras(precision=None, blocks=8):
    """The reference's E1 configuration (config_files/RDST_E1_OASIS_example_SRx4.ini as read by its own
    utils/param_loader.ParametersLoader; committed as tests/golden/e1_paras.json by oracle/gen_paras.py).
    blocks=4 gives the light "RDST-E" variant of BASELINE cfg3."""
    with open(os.path.join(ROOT, "tests", "golden", "e1_paras.json")) as f:
        p = types.SimpleNamespace(**json.load(f)["paras"])
    for name in ("rdst_dense_layer_depths", "rdst_num_heads", "rdst_window_size", "rdst_rdb_depths"):
        setattr(p, name, list(getattr(p, name))[:blocks])
    if precision is not None:
        p.rdst_b200_precision = precision
    return p


HR_PIX_PER_VOLUME = SLICES * LR_H * SCALE * LR_W * SCALE
FLOP_PER_LR_PX = 10_592_280                       # algorithmic, SURVEY 8(d)
ATTN_FLOP_PER_WINDOW = {60: 2_826_240, 90: 5_621_760, 120: 9_338_880}   # 512 C^2 + 16384 C


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, names, reasons = [], ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 6:
                    continue
                sm.append(float(f[0]))
                out["sm_max_mhz"] = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        return out


def _dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def _reference_module():
    """The reference's own RDSTSR built by its own factory, if /root/reference is importable here (build container);
    None on the GPU box.  Returns (module, kind)."""
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "networks")):
        return None
    try:
        for p_ in (os.path.join(ROOT, "oracle", "_shim"), ref):
            if p_ not in sys.path:
                sys.path.insert(0, p_)
        from networks.rdst_variations import make_RDSTSR as ref_make          # noqa: E402
        torch.manual_seed(0)
        return ref_make(e1_paras()).eval()
    except Exception as e:  # noqa: BLE001
        sys.stderr.write(f"bench.py: reference module not importable ({e!r}); using the oracle port\n")
        return None


def cpu_reference_run(sd, steps, warmup, sample_slices, module=None, budget_s=None):
    """Times the reference on the host cores: the real module when given, else the CPU restatement (oracle, PyTorch CPU
    fp32 -- the same aten ops the reference module executes).  With `budget_s`, one probe forward decides how many of
    the requested slices a step can hold so that warmup + steps stay inside the budget (the batch is halved until it
    fits; never below 8 slices).  Returns (Mpix/s, seconds per step, threads, slices per step)."""
    import rdst_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    xs = torch.rand(sample_slices, 1, LR_H, LR_W, generator=torch.Generator().manual_seed(1))
    fwd = (lambda x: module(x)) if module is not None else (lambda x: O.forward(sd, x, SCALE))
    n = sample_slices
    with torch.no_grad():
        if budget_s is not None:
            fwd(xs[:4])                                  # first-touch costs (thread pool, allocator) out of the probe
            t0 = time.perf_counter()
            fwd(xs[:8])
            per_slice = (time.perf_counter() - t0) / 8
            while n > 8 and per_slice * n * (steps + warmup) > budget_s:
                n = (n + 1) // 2
        x = xs[:n]
        for _ in range(warmup):
            fwd(x)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            fwd(x)
            ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    return n * LR_H * SCALE * LR_W * SCALE / sec / 1e6, sec, threads, n


_JSON_FD = None


def _protect_stdout():
    """Exactly ONE JSON line may reach stdout: libraries (NCCL prints its version banner there) get stderr instead."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the rdst_e_cfg3 / train_cfg4 extra keys")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    import rdst_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    config = {"workload": f"RDST-E1 x4 inference, {SLICES} synthetic OASIS LR slices 1x{LR_H}x{LR_W} per GPU per step "
                          f"(-> {SLICES}x{LR_H * SCALE}x{LR_W * SCALE} HR), random-init weights",
              "slices_per_gpu": SLICES, "lr_hw": [LR_H, LR_W], "sr_scale": SCALE,
              "parallelism": f"slice-sharded x{max(world, 1)} (no collective)", "precision": args.precision,
              "l2": "256 MiB buffer written between timed steps (L2 flush)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        ref_m = _reference_module()
        torch.manual_seed(0)
        sd = None
        if ref_m is None:
            m = rdst_b200.make_RDSTSR(e1_paras("fp32"))
            sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        steps, warm = max(1, args.steps), max(0, args.warmup)
        val, sec, threads, n_used = cpu_reference_run(sd, steps, warm, SLICES, ref_m,
                                                       budget_s=float(os.environ.get("RDST_BENCH_REF_BUDGET_S", "240")))
        kind = "reference" if ref_m is not None else "port"
        line = {"impl": "reference", "metric": "HR output Mpix/s (RDST-E1 x4 inference)", "value": round(val, 4),
                "unit": "Mpix/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
                "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": round(val, 4), "unit": "Mpix/s", "cores": threads, "kind": kind,
                                 "sample": f"{n_used} of {SLICES} slices per step, one forward of batch {n_used} ("
                                           + ("the reference module networks/rdst_variations.py built by its own make_RDSTSR"
                                              if kind == "reference" else
                                              "oracle = PyTorch-CPU restatement of the reference module; the Python "
                                              "reference itself cannot travel to the box") + ")"},
                "e2e": {"value": round(val, 4), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(line)
        return

    # ------------------------------------------------------------------ our arm (GPU)
    world, rank, local = _dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from rdst_b200 import executor as ex_mod
    torch.manual_seed(0)
    m = rdst_b200.make_RDSTSR(e1_paras(args.precision)).to(dev).eval()      # the reference-facing factory call
    x_host = torch.rand(SLICES, 1, LR_H, LR_W, generator=torch.Generator().manual_seed(1 + rank)).pin_memory()
    x_dev = x_host.to(dev)
    y_host = torch.empty(SLICES, 1, LR_H * SCALE, LR_W * SCALE).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # ---- parity gate: sampled slices of THIS batch against the CPU oracle (checker only; nothing here is timed)
    parity = None
    if rank == 0:
        import rdst_oracle as O
        with torch.no_grad():
            y_chk = m(x_dev).cpu()
            sd_cpu = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
            ref_chk = O.forward(sd_cpu, x_host[PARITY_SLICES].clone(), SCALE)
        parity = float((y_chk[PARITY_SLICES] - ref_chk).abs().max())
        tol = 1e-2 if args.precision == "bf16" else 1e-4
        if not (parity < tol):
            raise SystemExit(f"bench.py: parity gate failed: max|y - oracle| = {parity:.3e} >= {tol} on slices {PARITY_SLICES}")
        del y_chk

    launches = [0]
    ktime = {"events": [], "on": False, "hbm_events": []}
    TOP = "rdst_stl_attn_fwd_bf16"
    HBM_KERNEL = "rdst_last_conv_fwd_bf16_tc"        # the step's HBM-bound kernel (reads the 64-channel HR feature map once)
    orig_call = ex_mod.call

    def counting_call(name, *a):
        launches[0] += 1
        if ktime["on"] and name == TOP and a[12] == 120:
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            orig_call(name, *a)
            e1.record()
            ktime["events"].append((e0, e1, a[12]))          # a[12] = C of this launch
        elif ktime["on"] and name == HBM_KERNEL:
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            orig_call(name, *a)
            e1.record()
            ktime["hbm_events"].append((e0, e1))
        else:
            orig_call(name, *a)

    ex_mod.call = counting_call

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def run(step_fn, steps):
        evs = []
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            step_fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)          # ms over all steps (device time, flush excluded)

    def step_resident():
        with torch.no_grad():
            m(x_dev)

    def step_e2e():
        # the module's public call: pinned host input -> H2D copy -> forward; the HR volume lands in pinned host memory
        # (out=: the reconstruction kernel's stores cross PCIe as they are produced, inside the timed region)
        with torch.no_grad():
            xd = x_host.to(dev, non_blocking=True)
            m(xd, out=y_host)

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches[0] = 0
    ms_total = run(step_resident, args.steps)          # the headline timed region: no per-kernel events inside
    n_launch = launches[0]
    barrier()
    # same K steps again with CUDA-event pairs around the dominant kernel's launches (roofline numerator); kept out of
    # the headline region because an event between two launches defeats their programmatic-dependent-launch overlap
    ktime["on"] = args.precision == "bf16"
    ms_instr = run(step_resident, args.steps)
    ktime["on"] = False
    barrier()
    clocks = sampler.stop()
    for _ in range(2):
        step_e2e()
    barrier()
    ms_e2e = run(step_e2e, args.steps)
    barrier()

    # ---- extra keys (not the headline): RDST-E (cfg3 model) on the same batch, and the cfg4 training step
    extras = {}
    if not args.no_extras and args.precision == "bf16":
        extras = _extras(args, world, rank, dev, x_dev, run, barrier)

    t = torch.tensor([ms_total, ms_e2e] + [extras.get(k, 0.0) for k in ("_e_ms", "_train_ms", "_strong_ms")], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, e_ms, train_ms, strong_ms = (float(v) for v in t)
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = world * HR_PIX_PER_VOLUME / (ms_step * 1e-3) / 1e6
    e2e_val = world * HR_PIX_PER_VOLUME / (ms_e2e / args.steps * 1e-3) / 1e6
    tf_peak, hbm_peak, which = _peaks()
    roof = None
    if ktime["events"]:
        # dominant kernel: the fused window-attention kernel at C=120 (largest single share of the step)
        nwin = SLICES * (LR_H // 8) * (LR_W // 8)
        n_ev = len(ktime["events"])
        tot_ms = sum(a.elapsed_time(b) for a, b, _ in ktime["events"])
        avg_s = tot_ms * 1e-3 / n_ev
        ach = ATTN_FLOP_PER_WINDOW[120] * nwin / avg_s / 1e12
        alg_bytes = 2 * nwin * 64 * 128 * 2                  # read + write of [T][128] bf16 rows
        traffic = None
        for tf_name in ("r2_traffic.json", "r1_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", tf_name)) as f:
                    tj = json.load(f)["stl_attn_kernel<120>"]
                traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]
                break
            except Exception:  # noqa: BLE001
                pass
        roof = {"bound": "tensor", "kernel": "a2::stl_attn2_kernel<120> (fused LN+QKV+QK^T+softmax+PV+proj, one launch, warp-specialised)",
                "achieved": round(ach, 2), "peak": tf_peak, "unit": "TFLOP/s", "frac": round(ach / tf_peak, 4),
                "traffic": traffic, "peak_source": f"{which} (bf16_tflops_sustained)",
                "algorithmic_flop_per_launch": ATTN_FLOP_PER_WINDOW[120] * nwin, "algorithmic_bytes_per_launch": alg_bytes,
                "hbm_view": {"achieved_gbs": round(alg_bytes / avg_s / 1e9, 1), "peak_gbs": hbm_peak,
                             "frac": round(alg_bytes / avg_s / 1e9 / hbm_peak, 4)},
                "launches_timed": n_ev, "avg_launch_us": round(avg_s * 1e6, 2),
                "share_of_step": round(tot_ms / ms_instr, 4), "ms_per_step_instrumented": round(ms_instr / args.steps, 3)}
    hbm_roof = None
    if ktime["hbm_events"]:
        # reconstruction conv (64 -> 1 channels, 3x3) on the x4 feature map: algorithmic bytes = the bf16 input read once
        # (128 B per HR pixel) + the fp32 image written once (4 B per HR pixel)
        hr_px = SLICES * LR_H * SCALE * LR_W * SCALE
        bytes_alg = hr_px * (64 * 2 + 4)
        avg_s2 = sum(a.elapsed_time(b) for a, b in ktime["hbm_events"]) * 1e-3 / len(ktime["hbm_events"])
        hbm_roof = {"bound": "hbm", "kernel": "last_conv_tap_kernel (3x3 reconstruction conv as a tap GEMM, TMA-fed)",
                    "achieved": round(bytes_alg / avg_s2 / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(bytes_alg / avg_s2 / 1e9 / hbm_peak, 4), "algorithmic_bytes_per_launch": bytes_alg,
                    "avg_launch_us": round(avg_s2 * 1e6, 2), "launches_timed": len(ktime["hbm_events"]), "traffic": None}
        try:
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tj = json.load(f)["last_conv_tap_kernel"]
            hbm_roof["traffic"] = tj["dram_read_bytes"] + tj["dram_write_bytes"]
        except Exception:  # noqa: BLE001
            pass
    whole = FLOP_PER_LR_PX * SLICES * LR_H * LR_W * world / (ms_step * 1e-3) / 1e12
    line = {"metric": "HR output Mpix/s (RDST-E1 x4 inference)", "value": round(value, 2), "unit": "Mpix/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": n_launch,
            "parity_max_abs": parity, "parity_slices": PARITY_SLICES,
            "e2e": {"value": round(e2e_val, 2), "unit": "Mpix/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": y_host.numel() * 4},
            "whole_net_tflops": round(whole, 2), "whole_net_frac_of_tensor_peak": round(whole / tf_peak / world, 4)}
    if roof:
        line["roofline"] = roof
    if hbm_roof:
        line["roofline_hbm_kernel"] = hbm_roof
    if e_ms > 0:
        line["rdst_e_cfg3"] = {"what": "RDST-E (4 RDSTBs) x4 bf16, same 176-slice batch per GPU, inputs resident",
                               "ms_per_step": round(e_ms, 3),
                               "value": round(world * HR_PIX_PER_VOLUME / (e_ms * 1e-3) / 1e6, 2), "unit": "Mpix/s"}
    if strong_ms > 0:
        line["rdst_e_cfg3_strong"] = {"what": f"ONE {SLICES}-slice RDST-E volume split over {world} GPU(s) (contiguous slice shares, no "
                                              "collective), one CUDA-graph replay per rank, inputs resident, max over ranks",
                                      "scaling": "strong", "ms_per_volume": round(strong_ms, 3),
                                      "value": round(HR_PIX_PER_VOLUME / (strong_ms * 1e-3) / 1e6, 2), "unit": "Mpix/s"}
    if train_ms > 0:
        line["train_cfg4"] = dict(extras["_train_info"], ms_per_step=round(train_ms, 3),
                                  value=round(world * 32 * 96 * 96 / (train_ms * 1e-3) / 1e6, 3), unit="HR Mpix/s")
    if world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        val, sec, threads, n_used = cpu_reference_run(sd, 2, 1, 32, budget_s=25.0)
        line["cpu_baseline"] = {"value": round(val, 4), "unit": "Mpix/s", "cores": threads, "kind": "port",
                                "sample": f"{n_used} of {SLICES} slices per step, 2 steps, {sec:.2f} s/step "
                                          "(oracle = PyTorch-CPU restatement of the reference module)"}
    _emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def _extras(args, world, rank, dev, x_dev, run, barrier):
    """cfg3 model on the bench batch and the cfg4 training step (device time, ms per step, this rank)."""
    import rdst_b200
    out = {}
    torch.manual_seed(0)
    me = rdst_b200.make_RDSTSR(e1_paras("bf16", blocks=4)).to(dev).eval()

    def step_e():
        with torch.no_grad():
            me(x_dev)

    for _ in range(3):
        step_e()
    barrier()
    out["_e_ms"] = run(step_e, args.steps) / args.steps
    # ---- cfg3 strong scaling: ONE 176-slice RDST-E volume over all ranks (contiguous slice shares, no collective), each
    # rank replaying one CUDA graph of its share (rdst_b200.infer: super_resolve_volume_sharded(use_graph=True))
    from rdst_b200 import infer as rinfer
    b0, b1 = rinfer.shard_range(SLICES, world, rank)
    xs = x_dev[b0:b1].contiguous()
    ys = torch.empty(b1 - b0, 1, LR_H * SCALE, LR_W * SCALE, device=dev)

    def step_strong():
        rinfer.super_resolve_slices(me, xs, batch_size=SLICES, out=ys, use_graph=True)

    for _ in range(3):
        step_strong()
    barrier()
    out["_strong_ms"] = run(step_strong, args.steps) / args.steps
    del me
    # ---- cfg4: one data-parallel training step, whole step (fwd, L1, bwd, NCCL all-reduce, Adam) as one CUDA graph
    from rdst_b200 import ddp as rddp, train as rtrain
    torch.manual_seed(0)
    mt = rdst_b200.make_RDSTSR(e1_paras("bf16")).to(dev).train()
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    xb = torch.rand(32, 1, 24, 24, device=dev, generator=g)
    yb = torch.rand(32, 1, 96, 96, device=dev, generator=g)
    opt = torch.optim.Adam([p for p in mt.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99), eps=1e-8,
                           capturable=True, fused=True)
    reducer = rddp.BucketedAllReduce(mt) if world > 1 else None
    stepper = rtrain.GraphedTrainStep(mt, opt, xb, yb, reducer=reducer)
    losses = []
    for _ in range(3):
        losses.append(float(stepper(xb, yb)))
    barrier()
    steps = max(args.steps, 5)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        loss = stepper(xb, yb)
    e1.record()
    torch.cuda.synchronize()
    losses.append(float(loss))
    out["_train_ms"] = e0.elapsed_time(e1) / steps
    out["_train_info"] = {"what": "RDST-E1 x4 training step, 32 x 1x24x24 per GPU, L1 loss, Adam, bf16 tensor-core GEMMs, "
                                  "whole step one CUDA graph" + (", NCCL gradient all-reduce per link inside the graph"
                                                                 if world > 1 else ""),
                          "steps": steps, "loss_first": round(losses[0], 5), "loss_last": round(losses[-1], 5)}
    if reducer is not None:
        reducer.remove()
    barrier()
    return out


if __name__ == "__main__":
    main()
