"""Static check of the persistent kernels' tile loops: lists indexed constant loads (LDC c[0x3][R..]) and local
memory traffic (LDL / STL) that sit INSIDE a loop (between a backward branch and its target), per kernel.
usage: sass_check.py [lib.so] [name-substring]"""
import re
import subprocess
import sys
from pathlib import Path

so = Path(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else \
    Path(__file__).resolve().parents[1] / "ml_audio_inpainting_b200" / "lib" / "libaip_b200.so"
want = sys.argv[-1] if len(sys.argv) > 1 and not sys.argv[-1].endswith(".so") else "512"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = funcs.setdefault(m.group(1), [])
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur is not None:
        cur.append((int(m.group(1), 16), m.group(2)))
for name, ins in funcs.items():
    if want not in name:
        continue
    loops = []
    for addr, text in ins:
        m = re.search(r"\bBRA(?:\.\S+)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    big = [(a, b) for a, b in loops if b - a > 0x400]
    # the out-of-line mbarrier spin paths jump back INTO the loops from the end of the function: keep innermost only
    big = [l for l in big if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in big)]
    big = [l for l in big if not any(o != l and o[0] < l[0] < o[1] < l[1] for o in big)]
    bad = []
    for addr, text in ins:
        if re.search(r"LDC(\.64)? R\d+, c\[0x3\]\[R", text) or re.search(r"\b(LDL|STL)\b", text):
            inside = [l for l in big if l[0] <= addr <= l[1]]
            if inside:
                bad.append((addr, text.strip()))
    print(f"{name}: {len(ins)} instr, big loops {[(hex(a), hex(b)) for a, b in big]}, in-loop LDC/LDL/STL: {len(bad)}")
    for addr, text in bad[:12]:
        print(f"    {addr:#06x}  {text[:70]}")
