"""Print the SASS of one kernel between two addresses.  usage: sass_dump.py <obj> <mangled-substring> [lo hi]"""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
f = False
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m: f = pat in m.group(1); continue
    if not f: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        a = int(m.group(1), 16)
        if lo <= a <= hi: print("%04x %s" % (a, m.group(2).strip()))
