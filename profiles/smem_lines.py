"""Per-CUDA-source-line shared-memory wavefronts (total / excessive = bank conflicts) of one kernel from an .ncu-rep.
usage: python profiles/smem_lines.py <rep> <kernel-regex> [top_n]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
fname = None; hdr = None; agg = {}
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] in ("File Path", "File Name"): fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit() and len(r) >= len(hdr) - 2:
        def num(n):
            try: return float(r[hdr.index(n)])
            except ValueError: return 0.0
        a = agg.setdefault((fname, int(r[0])), [r[1].strip()[:100], 0, 0, 0, 0])
        a[1] += num("Instructions Executed"); a[2] += num("L1 Wavefronts Shared"); a[3] += num("L1 Wavefronts Shared Excessive"); a[4] += num("L1 Wavefronts Shared Ideal")
tw = sum(a[2] for a in agg.values()) or 1; tx = sum(a[3] for a in agg.values()); ti = sum(a[1] for a in agg.values()) or 1
print("shared-memory wavefronts %.4g, excessive %.4g (%.1f %%)  [all launches of the kernel in the report]" % (tw, tx, 100 * tx / tw))
print("wave%   exc%  wavefronts/ideal  inst%  file:line  source")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][3])[:top]:
    print("%5.1f  %5.1f  %6.2f  %5.1f  %s:%d  %s" % (100 * a[2] / tw, 100 * a[3] / tw, a[2] / max(a[2] - a[3], 1), 100 * a[1] / ti, k[0], k[1], a[0]))
