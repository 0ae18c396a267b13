"""Per-CUDA-source-line executed instructions / stall samples of one kernel in an ncu report (needs -lineinfo and
--import-source on).  usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top_n]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
lines = []
cur_file = ""
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or r[0] in ("File Path", "Function Name") or r[0] == "": continue
    try:
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ie]), int(r[isamp])))
    except ValueError:
        pass
tot = sum(l[3] for l in lines) or 1; tots = sum(l[4] for l in lines) or 1
print(f"total warp-inst {tot}  samples {tots}")
for f, n, s, e, sm in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{f}:{n:<4d} {100*e/tot:5.1f}% inst {100*sm/tots:5.1f}% smp  {s[:110]}")
