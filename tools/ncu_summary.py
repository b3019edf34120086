"""Summarise an .ncu-rep (ncu --set full) into the plain-text form kept under profiles/: one block per captured
launch with the counters DESIGN.md argues from.  Usage: python tools/ncu_summary.py REPORT.ncu-rep [kernel regex]"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0]
        if (pat and not pat.search(name)) or name in seen:
            continue
        seen.add(name)
        print("## %s" % name)
        for m in METRICS:
            if m in ix:
                print("%-70s %s %s" % (m, r[ix[m]], units[ix[m]]))
        print()


if __name__ == "__main__":
    main()
