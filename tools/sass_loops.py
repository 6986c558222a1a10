"""Histogram of SASS opcodes inside each loop (backward branch) of a kernel.
usage: sass_loops.py <obj-or-so> <mangled-substring>"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
FMA_PIPE = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL"}
cur = None; funcs = {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name: continue
    print("==", name, len(ins), "instructions")
    for addr, txt in ins:
        m = re.search(r"\bBRA\S*\s+(?:.*?)(0x[0-9a-f]+)", txt)
        if not m: continue
        tgt = int(m.group(1), 16)
        if tgt >= addr: continue
        body = [t for a, t in ins if tgt <= a <= addr]
        h = collections.Counter()
        for t in body:
            t = re.sub(r"^@!?U?P\w+\s+", "", t)
            op = t.split()[0]
            h[op.split(".")[0] + (".WIDE" if ".WIDE" in op else "")] += 1
        fma = sum(c for o, c in h.items() if o.split(".")[0] in FMA_PIPE)
        print(f"loop {tgt:#x}..{addr:#x}: {len(body)} instr, fma-pipe {fma}, other {len(body)-fma}")
        print("   ", dict(h.most_common()))
