"""Summary of an `ncu --set full` report in the format of profiles/*_ncu_full_summary.txt:
    ncu -i rep.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv [kernel-name-regex] > profiles/rNN_ncu_full_summary.txt
One block per distinct kernel (first launch that matches): the metrics bench.py and DESIGN.md quote."""
import csv
import re
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__t_sectors.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[start], rows[start + 1]
    name_col = hdr.index("Kernel Name")
    seen = set()
    for r in rows[start + 2:]:
        if len(r) != len(hdr):
            continue
        name = r[name_col]
        if name in seen or (pat and not pat.search(name)):
            continue
        seen.add(name)
        print("==== " + name)
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("   %s %s %s" % (m, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            mm = STALL.match(h)
            if mm:
                try:
                    stalls.append((float(r[i].replace(",", "")), mm.group(1) or mm.group(2)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print("   stalls: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
