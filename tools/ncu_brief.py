"""Brief per-kernel digest of an .ncu-rep (--set full): time, throughputs, occupancy, top stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("==", r[col["Kernel Name"]][:90])
    for k in keys:
        if k in col:
            print(f"   {k:82s} {r[col[k]]} {units[col[k]]}")
    st = []
    for h, i in col.items():
        if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued"):
            try:
                st.append((float(r[i].replace(",", "")), h.split("stalled_")[1]))
            except ValueError:
                pass
    tot = sum(v for v, _ in st) or 1
    print("   stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:7]))
