"""Summarise an .ncu-rep: python tools/ncu_summary.py file.ncu-rep [--source N]
Prints per-launch key metrics (raw page) and, with --source, the N hottest SASS lines by stall samples."""
import csv, subprocess, sys, io
rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__icc_requests.sum", "gcc__", "smsp__pcsamp_warps_issue_stalled"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==== ", r[hdr.index("Kernel Name")][:80], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
    for i, hname in enumerate(hdr):
        if any(hname.startswith(k) for k in KEYS) and r[i] not in ("", "0"):
            print("  %-90s %-10s %s" % (hname, units[i], r[i]))
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r))
    hdr = rows[hi]
    si = hdr.index("Source")
    samp = next(i for i, c in enumerate(hdr) if c.startswith("# Samples") or c.startswith("Warp Stall Sampling (All"))
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    tot = sum(float(r[samp] or 0) for r in body)
    print("total samples", tot, "sass lines", len(body))
    order = sorted(range(len(body)), key=lambda i: -float(body[i][samp] or 0))[:n]
    for i in sorted(order):
        print("%6d %8s  %s" % (i, body[i][samp], body[i][si][:110]))
