"""Turns gpurun_out/*.ncu-rep and launch-list CSVs into small text summaries kept under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.txt
  python profiles/summarize.py kernel   gpurun_out/prof_count_r1.ncu-rep > profiles/r1_count_kernel.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum", "lts__t_sectors_op_atom.sum",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    print(f"# {path}: gpu__time_duration.sum per kernel over one timed step (ncu --clock-control none; cold, serialised: compare shares)")
    for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{v:10.1f} us {100 * v / tot:5.1f}%  x{n:3d}  {k}")
    print(f"{tot:10.1f} us total")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full --clock-control none")
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")][:100])
        for i, h in enumerate(hdr):
            if h in KEYS:
                v, u = vals[i], units[i]
                if u in ("Gbyte", "Kbyte", "byte", "Tbyte"):  # normalise to Mbyte so summaries compare
                    v = f"{float(v.replace(',', '')) * {'Tbyte': 1e6, 'Gbyte': 1e3, 'Kbyte': 1e-3, 'byte': 1e-6}[u]:.6f}"
                    u = "Mbyte"
                print(f"  {h:75s} {v:>18s} {u}")
            elif "warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and vals[i] and float(vals[i]) > 3:
                print(f"  {h:75s} {vals[i]:>18s} %")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
