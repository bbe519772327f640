"""Per-kernel counts of the Blackwell-only SASS opcodes in librdst_b200.so (cuobjdump -sass): the evidence that the hot
kernels use tcgen05 (UTC*MMA), TMEM loads/stores (LDTM/STTM) and TMA (UTMALDG/UTMASTG/UBLKCP) -- and that none uses the
legacy mma.sync path (HMMA).      python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "rdst_b200", "lib", "librdst_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
ops = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS", "HMMA", "MUFU.EX2"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for o in ops:
            if re.search(r"\b" + re.escape(o) + r"\b", line) or (o + ".") in line:
                counts[cur][o] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a); kernels without any listed opcode omitted")
print("kernel".ljust(70) + "".join(o.rjust(9) for o in ops))
for k, c in counts.items():
    if sum(c.values()):
        print(k[:69].ljust(70) + "".join(str(c[o]).rjust(9) for o in ops))
