"""Pretty-print a role timeline file written by tools/attn2_timing.py --full (absolute cycles, CTA 0), with labels.
usage: a2_parse.py file lo hi [C]"""
import re, sys
f = open(sys.argv[1]).read().splitlines()
C = int(sys.argv[4]) if len(sys.argv) > 4 else 120
NHALF = 1 if C == 60 else 2
roles = {}
for l in f:
    m = re.match(r'role (\w) \((\d+) stamps\): (.*)', l)
    if m:
        roles[m.group(1)] = [int(v) for v in m.group(3).split()]
lo, hi = int(sys.argv[2]), int(sys.argv[3])
ev = []
def lab(r, i):
    if r == "C": return f"C {'S ready' if i % 2 == 0 else 'P written'} g={2 * (i // 2)}"
    if r == "D": return f"D {'S ready' if i % 2 == 0 else 'P written'} g={2 * (i // 2) + 1}"
    if r == "M":          # issuer of slot 0 (even heads): 3 stamps per iteration
        it, k = divmod(i, 3)
        j = 2 * it
        return f"M0 j={j}: " + ["P,V ready", "PV issued", f"S({j + 2}) (+qkv({j + 4})) issued"][k]
    if r == "B":
        per = 13
        t, k = divmod(i, per)
        # per tile: heads 0..5 -> (qkv ready, drained) ; the 'normalised O complete' stamp of the PREVIOUS tile comes in iteration g=6t+1 after its drain
        seq = []
        for h in range(6):
            seq += [f"qkv ready g={6*t+h}", f"qk drained g={6*t+h}"]
            if h == 1 and t >= 1: seq += [f"AP complete tile {t-1}"]
        if t == 0:
            per0 = 12
            if i < per0: return "B " + seq[i]
        # generic fallback
        return f"B stamp {i}"
    if r == "A": return f"A stamp {i}"
for r, st in roles.items():
    for i, t in enumerate(st):
        if lo <= t <= hi:
            ev.append((t, r, i))
ev.sort()
prev = None
for t, r, i in ev:
    print(f"{t:8d}  {'      ' * 'ABCDM'.index(r)}{lab(r, i)}")
