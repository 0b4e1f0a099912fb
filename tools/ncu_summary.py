"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into the handful of metrics DESIGN.md cites.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/prof_summary.txt"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum"]
STALL = "smsp__average_warps_issue_stalled_"

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("== %s\n   kernel: %s" % (rep, d.get("Kernel Name", "?")[:150]))
        for k in KEYS:
            if k in d:
                print("   %-70s %s %s" % (k, d[k], u[k]))
        st = sorted(((float(d[h]), h[len(STALL):-len("_per_issue_active.ratio")]) for h in hdr
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h]), reverse=True)
        print("   warp stalls per issue slot: " + ", ".join("%s %.2f" % (n, v) for v, n in st[:8]))
