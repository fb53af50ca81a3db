"""ncu --set full report -> per-kernel DRAM traffic / duration / pipe utilisation.

    python tools/ncu_traffic.py gpurun_out/r02_kernels.ncu-rep profiles/r02_kernel_traffic.json > profiles/r02_ncu_full_kernels.md

The JSON maps a short kernel name to {"dram_bytes_per_launch", "duration_us", ...}; bench.py reads `roofline.traffic`
from it (the capture must be of the same kernel variant the bench times).
"""
import csv, io, json, re, subprocess, sys

rep, out_json = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, key):
    if key not in col:
        return None
    try:
        v = float(r[col[key]].replace(",", ""))
    except ValueError:
        return None
    u = units[col[key]].lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3,
             "second": 1e6}.get(u, 1.0)
    return v * scale


agg = {}
print("| kernel | launches | duration us | dram read MB | dram write MB | dram % | tensor % | issue % | regs |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = re.sub(r"\(.*", "", re.sub(r"<.*", "", name)).split("::")[-1].strip()
    full = re.sub(r"\(.*", "", name).strip()
    d = {"dur": val(r, "gpu__time_duration.sum"), "rd": val(r, "dram__bytes_read.sum"), "wr": val(r, "dram__bytes_write.sum"),
         "dram": val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
         "tensor": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
         or val(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
         "issue": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), "regs": val(r, "launch__registers_per_thread")}
    templ = full.split("::")[-1].strip()                 # name with its template arguments: one entry per variant
    for key in {short, templ}:
        a = agg.setdefault(key, {"n": 0, "dur": 0.0, "rd": 0.0, "wr": 0.0, "full": full})
        a["n"] += 1
        a["dur"] += d["dur"] or 0.0
        a["rd"] += d["rd"] or 0.0
        a["wr"] += d["wr"] or 0.0
    f = lambda v, p=1: "-" if v is None else f"{v:.{p}f}"
    print(f"| `{full[:70]}` | 1 | {f(d['dur'])} | {f((d['rd'] or 0) / 1e6)} | {f((d['wr'] or 0) / 1e6)} | {f(d['dram'])} | "
          f"{f(d['tensor'])} | {f(d['issue'])} | {f(d['regs'], 0)} |")
js = {k: {"launches": a["n"], "dram_bytes_per_launch": (a["rd"] + a["wr"]) / a["n"], "dram_read_bytes_per_launch": a["rd"] / a["n"],
          "dram_write_bytes_per_launch": a["wr"] / a["n"], "duration_us": a["dur"] / a["n"], "kernel": a["full"]}
      for k, a in agg.items()}
json.dump(js, open(out_json, "w"), indent=1)
