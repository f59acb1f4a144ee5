"""Opcode mix of one kernel in lib/libaip_b200.so (static SASS counts).  usage: sass_mix.py <substring of mangled name> [n]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parents[1] / "ml_audio_inpainting_b200" / "lib" / "libaip_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
want = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur = None
mix = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
    if m and cur:
        ins = m.group(1).split()
        op = ins[1] if ins[0].startswith("@") else ins[0]
        mix[cur][op.split(".")[0]] += 1
for name, c in mix.items():
    if want in name:
        print(name, sum(c.values()), "instructions")
        for op, n in c.most_common(top):
            print("  %-8s %5d" % (op, n))
