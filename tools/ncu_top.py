"""Read an .ncu-rep here (no GPU): headline counters and the most-stalled instructions of the first kernel in it.
usage: ncu_top.py file.ncu-rep [n_top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__average_warps_issue_stalled", "dram__throughput.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) for w in want) and "per_second" not in h and "pct_of_peak_sustained_elapsed" not in h:
        print("%-90s %-14s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
col = {n: i for i, n in enumerate(hdr)}
tot = sum(int(r[col["Warp Stall Sampling (All Samples)"]]) for r in data)
print("total stall samples", tot, "instructions", len(data))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {r: sum(int(x[col[r]]) for x in data) for r in reasons}
print("by reason:", ", ".join("%s %d" % (k[6:], v) for k, v in sorted(agg.items(), key=lambda t: -t[1]) if v))
top = sorted(enumerate(data), key=lambda t: -int(t[1][col["Warp Stall Sampling (All Samples)"]]))[:ntop]
for i, r in sorted(top):
    rs = sorted(((int(r[col[x]]), x[6:]) for x in reasons), reverse=True)[:2]
    print("%5d %-64s %6s exec %-9s %s" % (i, r[col["Source"]].strip()[:64], r[col["Warp Stall Sampling (All Samples)"]],
                                          r[col["Instructions Executed"]], " ".join("%s:%d" % (b, a) for a, b in rs if a)))
