"""Instructions executed per CUDA source line of an ncu report (needs -lineinfo and --import-source on): the per-role
instruction budget of a warp-specialised kernel.  usage: ncu_by_line.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; acc = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No", "Kernel Name"): continue
    if r[0].isdigit() and len(r) > 8 and r[7].isdigit():
        k = (cur, int(r[0]))
        a = acc.setdefault(k, [0, 0, r[1]])
        a[0] += int(r[7]); a[1] += int(r[6]) if r[6].isdigit() else 0
tot = sum(a[0] for a in acc.values())
print("total warp instructions", tot)
byfile = {}
for (f, l), a in acc.items(): byfile[f] = byfile.get(f, 0) + a[0]
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1]): print(f"  {f:24s} {v:10d} {100*v/tot:5.1f}%")
for (f, l), a in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0]:9d} {100*a[0]/tot:5.1f}% smp {a[1]:5d}  {f}:{l}  {a[2][:110]}")
