"""Per-source-line instruction breakdown of one kernel from an ncu report captured with --import-source on (needs -lineinfo).
usage: ncu_lines.py report.ncu-rep kernel-regex [file-substring] [min_share]
Prints, per source line (CUDA view): warp instructions executed, share, average active lanes, stall samples."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
minshare = float(sys.argv[4]) if len(sys.argv) > 4 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; per = {}; first_fn = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name":
        if first_fn is None: first_fn = r[1]
        elif r[1] != first_fn and cur_file is None: break
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "" or not r[0].isdigit(): continue
    ie, te, ns = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    key = (cur_file.split("/")[-1], int(r[0]))
    if key in per: continue  # a second kernel instance repeats the lines
    try:
        per[key] = (int(r[ie]), int(r[te]), int(r[ns]), r[1].strip())
    except ValueError:
        continue
tot = sum(v[0] for v in per.values()); ts = sum(v[2] for v in per.values())
print(f"kernel {first_fn[:90]}\ntotal warp instructions {tot}  samples {ts}")
for (f, ln), v in sorted(per.items()):
    if fsub in f and v[0] >= tot * minshare:
        print(f"{f}:{ln:4d} {v[0]/1e6:9.2f}M {100*v[0]/tot:5.1f}% lanes={v[1]/max(v[0],1):4.1f} samp%={100*v[2]/max(ts,1):5.1f}  {v[3][:110]}")

# optional grouping: env NCU_GROUPS="name:lo-hi,name:lo-hi" sums the lines of file `fsub` per range (intrinsic headers -> "inlined")
import os
if os.environ.get("NCU_GROUPS"):
    groups = [(g.split(":")[0], int(g.split(":")[1].split("-")[0]), int(g.split(":")[1].split("-")[1])) for g in os.environ["NCU_GROUPS"].split(",")]
    acc = {g[0]: [0, 0, 0] for g in groups}; acc["other"] = [0, 0, 0]
    for (f, ln), v in per.items():
        name = "other"
        if fsub and fsub in f:
            for g in groups:
                if g[1] <= ln <= g[2]: name = g[0]
        acc[name][0] += v[0]; acc[name][1] += v[1]; acc[name][2] += v[2]
    for k, v in acc.items():
        print(f"{k:12s} {v[0]/1e6:9.1f}M {100*v[0]/tot:5.1f}% lanes={v[1]/max(v[0],1):4.1f} samp%={100*v[2]/max(ts,1):5.1f}")
