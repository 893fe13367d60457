"""Headline counters of one kernel from an ncu report: duration, instructions, issue utilisation, stall reasons per issue, LSU pipe.
usage: ncu_stalls.py report.ncu-rep [row]"""
import csv, re, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h = rows[0]; row = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pat = r'issue_stalled_.*_per_issue_active.ratio|smsp__issue_active.avg.pct|gpu__time_duration.sum|smsp__inst_executed.sum$|l1tex__data_pipe_lsu_wavefronts.avg.pct|warps_active.avg.pct|dram__bytes_(read|write).sum$|launch__registers_per_thread$|smsp__sass_average_data_bytes_per_sector_mem_global_op_st'
print(rows[row][h.index("Kernel Name")][:80])
for i, k in enumerate(h):
    if re.search(pat, k):
        try:
            if float(rows[row][i]) > 0.3: print(f"  {k[:95]:95s} {rows[row][i]} {rows[1][i]}")
        except ValueError: pass
