"""Static SASS opcode histograms of the shipped library's kernels (cuobjdump -sass qblas_b200/libqblas_b200.so): the evidence that the
tensor kernel is a tcgen05 / TMA / TMEM kernel (UTCIMMA = tcgen05.mma kind::i8, UTMALDG = cp.async.bulk.tensor, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier) and what the integer kernels are made of.  usage: python tools/sass_digest.py > profiles/sass_digest.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "qblas_b200", "libqblas_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = [("k_oz_mmaILi1E", "k_oz_mma<1> (residue scheme tensor kernel)"), ("k_oz_mmaILi0E", "k_oz_mma<0> (plain int8 GEMM: microbenchmark / kernel test)"),
        ("k_crt_residuesILi5E", "k_crt_residues<5> (A rows, 5 words)"), ("k_crt_residues_tILi5E", "k_crt_residues_t<5> (B columns, transposing)"),
        ("k_crt_foldILi11E", "k_crt_fold<11> (41-44 moduli)"), ("k_crt_foldILi10E", "k_crt_fold<10> (37-40 moduli)"), ("k_crt_fixup", "k_crt_fixup"),
        ("k_oz_scan", "k_oz_scan"), ("k_gemv_f64ILb0E", "k_gemv_f64<row-major> (sliced FP64 qgemv: TMA tiles, DFMA)"), ("k_gemv_f64ILb1E", "k_gemv_f64<col-major>"),
        ("k_sumsq_tma", "k_sumsq_tma (sliced FP64 sum of squares fed by cp.async.bulk)"), ("k_gemv_row_wide", "k_gemv_row_wide (window accumulate, first instance)"), ("6k_gemm", "k_gemm (reference-order integer-limb qgemm, first version)"), ("9k_gemm_nb", "k_gemm_nb (reference-order qgemm, branch-free step: the default)"),
        ("k_dot_wide_tmaILi128ELi4ELb0E", "k_dot_wide_tma (fast qdot of two contiguous vectors: cp.async.bulk tiles, window accumulate)")]
funcs = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if cur and m:
        t = re.sub(r"^@!?U?P\w+\s+", "", m.group(1).strip()); op = t.split()[0]
        k = op.split(".")[0]
        if k in ("IMAD", "UTMALDG", "LDTM", "UTCIMMA", "UTCBAR", "SYNCS", "STG", "LDG", "IDP", "UBLKCP", "I2F"):
            k = ".".join(op.split(".")[:2])
        funcs[cur][k] += 1
print(f"# static SASS opcode counts, sm_100a, {os.path.relpath(so, ROOT)} (tools/sass_digest.py)")
seen = set()
for pat, title in want:
    for name, h in funcs.items():
        if pat in name and name not in seen:
            seen.add(name)
            tot = sum(h.values())
            print(f"\n== {title}: {name}\n   {tot} instructions")
            key = [k for k in h if k.split(".")[0] in ("UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "UTMAPF", "UTCATOMSWS", "FENCE", "IDP", "UBLKCP", "DFMA")]
            if key:
                print("   tensor / TMA / TMEM / mbarrier / dp4a: " + ", ".join(f"{k} x{h[k]}" for k in sorted(key)))
            print("   " + ", ".join(f"{k} {c}" for k, c in h.most_common(24)))
            break
