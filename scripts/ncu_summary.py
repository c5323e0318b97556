#!/usr/bin/env python
"""Turns the ncu artefacts of a gpurun call into the text summaries kept under profiles/.

  python scripts/ncu_summary.py full    <report.ncu-rep>  > profiles/rNN_step_hmm_ncu_full_summary.txt
  python scripts/ncu_summary.py launches <launches.csv>   > profiles/rNN_launch_list_summary.txt

`full`: selected raw-page metrics of the first kernel in the report + warp-stall samples summed over the
source page + instructions and stall samples per barrier interval of the SASS (one line per BAR.SYNC / mbarrier wait).
`launches`: per-kernel launch count, total and average duration, share.
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_shared_atom.sum",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def full(rep):
    rows = ncu(rep, "raw")
    h, units, r = rows[0], rows[1], rows[2]
    print("Kernel Name".ljust(80), r[h.index("Kernel Name")])
    for m in METRICS:
        if m in h:
            print(m.ljust(80), r[h.index(m)], units[h.index(m)])
    rows = ncu(rep, "source")
    hdr, data = rows[1], rows[2:]
    ix = {c: i for i, c in enumerate(hdr)}
    stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot = {c: sum(int(r[ix[c]]) for r in data) for c in stalls}
    T = sum(tot.values())
    print("\nwarp stall samples (source page, all instructions):")
    for c, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-28s %6d %5.1f%%" % (c, v, 100.0 * v / T))
    ntile = max(int(r[ix["Instructions Executed"]]) for r in data if "BAR.SYNC" in r[ix["Source"]])
    total = sum(int(r[ix["Instructions Executed"]]) for r in data)
    print("\nwarp instructions: %d total, %.1f per warp and children tile (%d warp-tiles)" % (total, total / ntile, ntile))
    print("per synchronisation interval of the SASS (linear address order; a barrier's wait samples land on the")
    print("instruction AFTER it, i.e. in the next interval):")
    acc = samp = 0
    st = dict.fromkeys(stalls, 0)
    for r in data:
        s = r[ix["Source"]]
        n = int(r[ix["Instructions Executed"]])
        acc += n
        samp += int(r[ix["# Samples"]])
        for c in stalls:
            st[c] += int(r[ix[c]])
        if "BAR.SYNC" in s or ("SYNCS" in s and "TRYWAIT" in s):
            if acc > 0.5 * ntile or samp > 20:
                top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
                print("%7.1f inst/warp-tile  %5d samples  ends with %-34s x%.2f | %s" % (
                    acc / ntile, samp, s.strip()[:34], n / ntile, ", ".join("%s=%d" % (k[6:], v) for k, v in top)))
            acc = samp = 0
            st = dict.fromkeys(stalls, 0)


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ix = {c: i for i, c in enumerate(rows[h])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[h + 2:]:
        if len(r) < len(rows[h]):
            continue
        k = r[ix["Kernel Name"]].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += float(r[ix["Metric Value"]].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("%-48s %8s %14s %8s %12s" % ("kernel", "launches", "total ns", "share", "avg ns"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-48s %8d %14.1f %7.1f%% %12.2f" % (k, v[0], v[1], 100.0 * v[1] / tot, v[1] / v[0]))


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
