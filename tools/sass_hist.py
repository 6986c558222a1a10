"""Opcode histogram of a SASS address range (dynamic count of a straight-line path).
usage: sass_hist.py <obj> <mangled-substring> <lo> <hi> [elements]"""
import re, subprocess, sys, collections
obj, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
per = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
f = False; h = collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m: f = pat in m.group(1); continue
    if not f: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if not m: continue
    a = int(m.group(1), 16)
    if not (lo <= a <= hi): continue
    t = re.sub(r"^@!?U?P\w+\s+", "", m.group(2).strip()); op = t.split()[0]; k = op.split(".")[0]
    if k == "IMAD":
        for tag in ("WIDE", "MOV", "HI", "SHL", "IADD"):
            if tag in op: k = "IMAD." + tag; break
        else:
            if op.endswith(".X") or ".X." in op: k = "IMAD.X"
    h[k] += 1
tot = sum(h.values())
fma = sum(c for k, c in h.items() if k.startswith("IMAD"))
lsu = sum(c for k, c in h.items() if k in ("LDS", "STS", "LDG", "STG", "LDL", "STL"))
uni = sum(c for k, c in h.items() if k.startswith("U") or k in ("S2UR", "R2UR"))
print(f"{tot} instr ({tot/per:.1f}/elem): fma-pipe {fma/per:.1f}, lsu {lsu/per:.1f}, uniform {uni/per:.1f}, alu+other {(tot-fma-lsu-uni)/per:.1f}")
print(dict(h.most_common()))
