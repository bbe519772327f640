"""Summarise an ncu report: headline metrics + stall breakdown + hottest SASS lines (needs -lineinfo, --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which captured launch the source page is read for
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or h == "Kernel Name":
        print(f"{h} [{units[i]}]:", [r[i][:60] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = idx[kidx]; end = idx[kidx + 1] - 1 if len(idx) > kidx + 1 else len(rows)
print(f"source-level samples of captured launch #{kidx}")
hdr = rows[start]; data = [r for r in rows[start + 1:end] if len(r) >= len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ci[s]]) for r in data) for s in stalls}
print("total samples", tot, "instructions", len(data))
for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {s:26s} {v:7d} {100 * v / max(tot,1):5.1f}%")
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:top_n]:
    st = max(stalls, key=lambda s: int(r[ci[s]]))
    print(r[ci["# Samples"]].rjust(6), r[ci["Instructions Executed"]].rjust(8), st[6:22].ljust(16), r[ci["Source"]][:100])
