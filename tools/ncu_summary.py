"""Summarise an ncu report (read on the CPU box): per kernel duration, DRAM bytes, throughputs.
usage: python tools/ncu_summary.py <report.ncu-rep> <out.json>"""
import csv, json, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
ki = hdr.index("Kernel Name")
res = []
for r in data:
    d = {"kernel": r[ki].split("(")[0].replace("vstabk::<unnamed>::", "")}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            d[w] = f"{r[i]} {units[i]}".strip()
    res.append(d)
json.dump(res, open(out, "w"), indent=1)
for d in res:
    print(d["kernel"][:40], d.get("gpu__time_duration.sum"), d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"),
          d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed"), d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
