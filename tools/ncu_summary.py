#!/usr/bin/env python
"""Text summary of an .ncu-rep (one kernel): the metrics DESIGN.md / profiles/README.md quote.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep [--hot]   > profiles/x.txt

--hot adds the per-instruction stall samples of the hottest loop (needs --import-source on at capture time).
"""
import csv
import subprocess
import sys


def main():
    fn = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", fn, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_tensor.sum",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
            "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    print("# %s" % fn)
    for k in keys:
        if k in d:
            print("%-70s %s %s" % (k, d[k], u[k]))
    print("# warp stall reasons (cycles per issued instruction)")
    st = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            st.append((float(d[h]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    for v, name in sorted(st, reverse=True):
        if v >= 0.01:
            print("  %-24s %.3f" % (name, v))
    if "--hot" in sys.argv:
        src = subprocess.run(["ncu", "-i", fn, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        h = rows[1]
        data = rows[2:]
        iS, iE, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
        ex = [int(r[iE]) for r in data]
        mx = max(ex)
        hot = [k for k, e in enumerate(ex) if e > 0.5 * mx]
        tot = sum(int(r[iS]) for r in data)
        hs = sum(int(data[k][iS]) for k in hot)
        print("# hottest loop: %d instructions executed %d times, %.1f %% of all stall samples" % (len(hot), mx, 100.0 * hs / tot))
        print("# index  samples  share  instruction")
        for k in hot:
            print("%5d %8d %5.2f%%  %s" % (k, int(data[k][iS]), 100.0 * int(data[k][iS]) / max(hs, 1), data[k][iSrc].strip()[:80]))


if __name__ == "__main__":
    main()
