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

def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(val.replace(",", "")) * mult


if len(sys.argv) > 1 and sys.argv[1] == "--traffic":
    # python tools/ncu_summary.py --traffic "kernel key=report.ncu-rep" ... > profiles/ncu_traffic_r01.json
    import json
    res = {}
    for arg in sys.argv[2:]:
        key, rep = arg.rsplit("=", 1)
        hdr, units, rows = raw_rows(rep)
        d, u = dict(zip(hdr, rows[0])), dict(zip(hdr, units))
        res[key] = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + \
            to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
    print(json.dumps(res, indent=1))
    sys.exit(0)

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
