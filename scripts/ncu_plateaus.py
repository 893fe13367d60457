"""Summarise an ncu source page (SASS view): runs of instructions with the same execution count.
usage: ncu_plateaus.py report.ncu-rep kernel-regex [min_share]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.005
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ie, src, ns, te = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples"), h.index("Thread Instructions Executed")
data = []
seen = set()
for r in rows[2:]:
    try:
        if r[0] in seen: break  # a second kernel instance repeats the addresses
        seen.add(r[0])
        data.append((r[0], r[src], int(r[ie]), int(r[ns]), int(r[te])))
    except Exception:
        pass
tot = sum(d[2] for d in data); ts = sum(d[3] for d in data)
print("total warp instr", tot, "samples", ts, "SASS instrs", len(data))
prev = None; start = 0; acc = accs = acct = 0; blocks = []
for i, d in enumerate(data):
    if prev is None or abs(d[2] - prev) > 0.03 * max(prev, 1):
        if prev is not None: blocks.append((start, i - 1, prev, acc, accs, acct))
        start = i; acc = accs = acct = 0
    acc += d[2]; accs += d[3]; acct += d[4]; prev = d[2]
blocks.append((start, len(data) - 1, prev, acc, accs, acct))
for b in blocks:
    if b[3] > tot * minshare:
        ops = " ".join(x[1].split()[0] if not x[1].startswith("@") else x[1].split()[1] for x in data[b[0]:min(b[1] + 1, b[0] + 7)])
        print(f"{b[0]:5d}-{b[1]:5d} n={b[1]-b[0]+1:4d} exec={b[2]/1e6:8.2f}M instr%={100*b[3]/tot:5.1f} samp%={100*b[4]/max(ts,1):5.1f} lanes={b[5]/max(b[3],1):4.1f}  {ops}")
