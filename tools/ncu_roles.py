"""Split an ncu source-page CSV of a warp-specialised kernel by address range (role) and report, per role,
executed warp-instructions, stall-sample shares and the opcode mix.
usage: ncu -i rep --page source --csv --kernel-name regex:... --launch-count 1 > src.csv
       python tools/ncu_roles.py src.csv name=0xLO-0xHI [name=0xLO-0xHI ...]   (offsets from the kernel start)"""
import csv
import sys
from collections import Counter, defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows[2:]:
    if len(r) < 10 or r[0] in seen or not r[0].startswith("0x"):      # repeated headers: the report holds several launches
        continue
    seen.add(r[0])
    data.append(r)
base = int(data[0][0], 16)
roles = []
for a in sys.argv[2:]:
    name, rng = a.split("=")
    lo, hi = rng.split("-")
    roles.append((name, int(lo, 16), int(hi, 16)))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
f = lambda r, h: float(r[idx[h]].replace(",", "") or 0)
tot_samples = sum(f(r, "# Samples") for r in data)
for name, lo, hi in roles:
    sel = [r for r in data if lo <= int(r[0], 16) - base <= hi]
    inst = sum(f(r, "Instructions Executed") for r in sel)
    smp = sum(f(r, "# Samples") for r in sel)
    st = {h: sum(f(r, h) for r in sel) for h in stall_cols}
    ops = Counter()
    for r in sel:
        t = r[idx["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += f(r, "Instructions Executed")
    print(f"== {name}: {len(sel)} SASS lines, {inst:.4g} warp-instr, {smp:.0f} samples ({100 * smp / tot_samples:.1f}% of all)")
    smp = smp or 1.0
    print("   stalls: " + ", ".join(f"{h[6:]}={100 * v / smp:.1f}%" for h, v in sorted(st.items(), key=lambda t: -t[1])[:9]))
    print("   opcodes: " + ", ".join(f"{o}={100 * n / inst:.1f}%" for o, n in ops.most_common(14)))
