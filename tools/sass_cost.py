"""Dispatch-cost estimate of a SASS address range: sum over the instructions of the clocks a warp instruction occupies one SM
sub-partition of sm_100a (tools/exp/mb_pipes.cu: IMAD.WIDE / IMAD.HI 4.3, IMAD / LOP3 / SHF / PRMT / LDS / STS 2, IADD3 / ISETP / SEL /
VIADD / VIMNMX / MOV 1).  The sum over the hot path of k_gemm reproduces its measured clocks per qFMA.
usage: sass_cost.py <obj-or-so> <mangled-substring> [lo hi]   (hex addresses; without them: every backward-branch loop)"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
COST = {"IMAD.WIDE": 4.3, "IMAD.HI": 4.3, "IMAD": 2, "LOP3": 2, "SHF": 2, "PRMT": 2, "LDS": 2, "STS": 2, "FLO": 2, "BREV": 2, "POPC": 2, "LEA": 2,
        "IADD3": 1, "ISETP": 1, "SEL": 1, "VIADD": 1, "VIMNMX": 1, "VIADDMNMX": 1, "MOV": 1, "CS2R": 1, "PLOP3": 1, "BRA": 1, "BSSY": 1, "BSYNC": 1,
        "IABS": 1, "LOP": 2, "R2P": 2, "P2R": 2}
cur = None; funcs = {}
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
def key(t):
    t = re.sub(r"^@!?U?P\w+\s+", "", t)
    op = t.split()[0]
    base = op.split(".")[0]
    if base == "IMAD" and (".WIDE" in op): return "IMAD.WIDE"
    if base == "IMAD" and (".HI" in op): return "IMAD.HI"
    return base
def report(body, title):
    h = collections.Counter(key(t) for _, t in body)
    cost = sum(COST.get(k, 1) * c for k, c in h.items())
    print(f"{title}: {len(body)} instr, est. {cost:.0f} dispatch clk")
    print("   ", dict(h.most_common()))
for name, ins in funcs.items():
    if pat not in name: continue
    print("==", name, len(ins), "instructions")
    if len(sys.argv) >= 5:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
        report([(a, t) for a, t in ins if lo <= a <= hi], f"range {lo:#x}..{hi:#x}")
        continue
    for addr, txt in ins:
        m = re.search(r"\bBRA\S*\s+(?:.*?)(0x[0-9a-f]+)", txt)
        if not m: continue
        tgt = int(m.group(1), 16)
        if tgt >= addr: continue
        report([(a, t) for a, t in ins if tgt <= a <= addr], f"loop {tgt:#x}..{addr:#x}")
