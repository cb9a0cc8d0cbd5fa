"""ptxas -v lines and the SASS mnemonics that show what the hot kernels are made of (TMA bulk copies, mbarrier waits,
cp.async, match.any, shared-memory atomics, 128-bit accesses), from the objects build.py leaves under
zlib.kotlin_b200/build/.  usage: sass_summary.py > profiles/sass_r2.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "zlib.kotlin_b200", "build")
KERNELS = {"zk_lz77": ["lz77_kernel"], "zk_huffman": ["hist_kernel", "huffman_kernel", "plan_offsets_kernel"], "zk_pack": ["pack_kernel", "stored_kernel"],
           "zk_adler": ["adler_spans_kernel"], "zk_inflate": ["decode_lanes_kernel", "expand_blocks_kernel", "block_find_kernel", "marker_compact_kernel", "materialize_kernel"]}
PATTERNS = ["UBLKCP", "SYNCS", "LDGSTS", "MATCH.ANY", "ATOMS", "RED.", "VOTE", "SHFL", "IDP", "LDS.128", "STS.128", "LDG.E.128", "STG.E.128", "BAR.SYNC", "POPC", "FLO"]
print("# SASS and ptxas facts of the hot kernels (round 2)\n")
print("`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xptxas -v`; counts are static SASS instructions per kernel")
print("(`cuobjdump -sass`), not executed ones.  Regenerate with `python tools/sass_summary.py`.\n")
for obj, kernels in KERNELS.items():
    o = os.path.join(B, obj + ".o")
    log = open(o + ".log").read()
    sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    arch = re.search(r"arch = (sm_\w+)", sass)
    funcs = re.split(r"\n\s*Function : ", sass)
    for k in kernels:
        body = next((f for f in funcs if k in f.split("\n", 1)[0]), None)
        if body is None:
            continue
        m = re.search(r"Compiling entry function '([^']*%s[^']*)'.*?\n(.*?)\n(ptxas info\s*: Used [^\n]*)" % k, log, re.S)
        print("## `%s` (%s.cu, %s)\n" % (k, obj, arch.group(1) if arch else "?"))
        if m:
            print("    " + m.group(2).strip())
            print("    " + m.group(3).strip())
        n = len(re.findall(r"/\*[0-9a-f]{4}\*/", body))
        counts = {p: len(re.findall(r"\b" + re.escape(p), body)) for p in PATTERNS}
        print("\n%d SASS instructions; " % n + ", ".join("%s %d" % (p, c) for p, c in counts.items() if c) + "\n")
        for p in ("UBLKCP", "SYNCS", "LDGSTS", "MATCH.ANY", "ATOMS.MAX", "RED"):
            ex = re.search(r"/\*[0-9a-f]{4}\*/\s+([^;]*\b%s[^;]*;)" % re.escape(p), body)
            if ex:
                print("    " + ex.group(1).strip())
        print()
