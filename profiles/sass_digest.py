"""SASS digest of libreflexiv_cuda.so: per kernel, the instruction count and the mnemonics that show HOW it runs
(TMA bulk copies, mbarrier ops, shared / global atomics, 128-bit loads).  `python profiles/sass_digest.py > profiles/r2_sass_digest.txt`"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "reflexiv_b200", "libreflexiv_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
watch = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "ATOM", "RED", "LDG.E.128", "STG.E.128", "LDS", "STS", "BAR", "SHFL", "MATCH", "VOTE", "BREV", "POPC", "LDC", "MEMBAR", "CCTL", "NANOSLEEP"]
kern = None
stats = {}
archs = set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        stats[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        archs.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        stats[kern]["_total"] += 1
        for w in watch:
            if op.startswith(w):
                stats[kern][w] += 1
                break
        if op.startswith("ATOMS.CAS") or op.startswith("ATOMS.CAST"):
            stats[kern]["ATOMS.CAS*"] += 1
        if ".64" in op and op.startswith("ATOMS"):
            stats[kern]["ATOMS.*.64"] += 1
print(f"# {os.path.basename(lib)}: cubins for {sorted(archs)}; {len(stats)} kernels")
print("# kernel | SASS instructions | watched mnemonics")
for k in sorted(stats, key=lambda k: -stats[k]["_total"]):
    c = stats[k]
    rest = ", ".join(f"{w} {c[w]}" for w in list(watch) + ["ATOMS.CAS*", "ATOMS.*.64"] if c[w])
    print(f"{k} | {c['_total']} | {rest}")
