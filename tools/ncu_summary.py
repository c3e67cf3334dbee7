"""Text summaries of ncu captures for profiles/ (development tool).

    python tools/ncu_summary.py launches gpurun_out/launches.csv           > profiles/rNN_bench_launches_summary.txt
    python tools/ncu_summary.py report   gpurun_out/x.ncu-rep ["note"]     > profiles/rNN_ncu_x.txt
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.max",
        "lts__t_sectors_srcunit_tex_op_red.sum"]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    iN, iV, iG = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
    agg = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= iV or "gpu__time_duration" not in ",".join(r):
            continue
        k = (r[iN][:78], r[iG])
        v = float(r[iV].replace(",", ""))
        unit = r[h.index("Metric Unit")]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    for (name, grid), (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-80s grid %-18s n=%4d mean_us %9.1f total_us %10.1f" % (name, grid, n, tot / n, tot))


def report(path, note=""):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    print("# " + path.split("/")[-1])
    if note:
        print("# " + note)
    for vals in rows[2:]:
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print("%-78s %s %s" % (k, vals[i], units[i]))
        stalls = [(float(vals[i] or 0), n) for i, n in enumerate(h) if "warp_issue_stalled" in n and n.endswith("per_warp_active.pct")]
        for v, n in sorted(stalls, reverse=True)[:6]:
            print("%-78s %.1f %%" % (n, v))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
