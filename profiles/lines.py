"""Per-CUDA-source-line share of executed warp instructions and stall samples from an .ncu-rep.
usage: python profiles/lines.py <rep> <kernel-regex> [min_pct]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[0].isdigit():
        ii = hdr.index("Instructions Executed"); wi = hdr.index("# Samples"); bi = hdr.index("stall_barrier")
        def num(x):
            try: return int(x)
            except: return 0
        agg.append((fname, int(r[0]), r[1], num(r[ii]), num(r[wi]), num(r[bi]), num(r[hdr.index("stall_long_sb")]), num(r[hdr.index("stall_short_sb")])))
ti = sum(a[3] for a in agg) or 1; ts = sum(a[4] for a in agg) or 1
# optional line-range buckets: name=file:lo-hi,... in env BUCKETS
import os
if os.environ.get("BUCKETS"):
    for spec in os.environ["BUCKETS"].split(","):
        name, rng = spec.split("=")
        f, lr = rng.split(":")
        lo, hi = map(int, lr.split("-"))
        bi = sum(a[3] for a in agg if a[0] == f and lo <= a[1] <= hi)
        bs = sum(a[4] for a in agg if a[0] == f and lo <= a[1] <= hi)
        print("bucket %-12s inst %5.1f%%  samples %5.1f%%" % (name, 100 * bi / ti, 100 * bs / ts))
    for f in sorted(set(a[0] for a in agg)):
        print("file %-28s inst %5.1f%%  samples %5.1f%%" % (f, 100 * sum(a[3] for a in agg if a[0] == f) / ti, 100 * sum(a[4] for a in agg if a[0] == f) / ts))
print("total warp instr %d, samples %d" % (ti, ts))
print(" inst%  samp%  (bar  lsb  ssb)  file:line  source")
for f, ln, src, i, s, b, l, sh in agg:
    if 100 * i / ti >= minp or 100 * s / ts >= minp:
        print("%5.1f  %5.1f  (%4.1f %4.1f %4.1f)  %s:%d  %s" % (100 * i / ti, 100 * s / ts, 100 * b / ts, 100 * l / ts, 100 * sh / ts, f, ln, src.strip()[:110]))
