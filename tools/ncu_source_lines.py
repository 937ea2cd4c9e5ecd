"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report captured with
--set full --import-source on:   python tools/ncu_source_lines.py report.ncu-rep [top_n]  >  profiles/<name>_source_lines.txt"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
agg = {}; tot = tots = 0; cur = None; hdr = None; kernel = None
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": kernel = r[1][:120]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr and len(r) > ie and r[0] != "":
        try: n = int(r[ie]); smp = int(r[isamp])
        except ValueError: continue
        a = agg.get((cur, int(r[0])), (0, 0, "")); agg[(cur, int(r[0]))] = (a[0] + n, a[1] + smp, r[1].strip()[:110]); tot += n; tots += smp
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
want = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
print("kernel:", kernel)
if len(rows) >= 3:
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if h in want: print(f"  {h} = {v} {u}")
print(f"warp instructions (source page): {tot}   stall samples: {tots}")
print(f"{'file:line':<28}{'instr %':>8}{'samples %':>10}  source")
for (f, l), (n, smp, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f + ':' + str(l):<28}{n / tot * 100:8.1f}{smp / max(tots, 1) * 100:10.1f}  {src}")
