"""Summarise an ncu capture exported as CSV: raw page (key counters, stalls per issue) and source page
(samples / executed instructions per 1 KB of code, hottest instructions).
   usage: ncu_read.py raw.csv source.csv [chunks]"""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
chunks = float(sys.argv[3]) if len(sys.argv) > 3 else 26421.0
rows = list(csv.reader(open(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want: print(f"{h:75s} {r[i]} {units[i]}")
st = []
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h and float(r[i]) > 0.05:
        st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
print("stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(rows[2][ia], 16)
data = [(int(x[ia], 16) - base, x[isrc].strip(), int(x[isamp]), int(x[iex])) for x in rows[2:]]
ts, te = sum(d[2] for d in data), sum(d[3] for d in data)
print(f"samples {ts}, warp instructions {te} = {te / chunks:.0f} per chunk")
cur, s_, e_ = 0, 0, 0
for a, t, sm, ex in data + [(1 << 30, "", 0, 0)]:
    if a // 0x400 != cur:
        if e_: print(f"  {cur * 0x400:6x}: samples {s_ / ts:6.1%}  instructions {e_ / chunks:7.0f} per chunk")
        cur, s_, e_ = a // 0x400, 0, 0
    s_ += sm; e_ += ex
for a, t, sm, ex in data:
    if sm >= ts * 0.008: print(f"  {a:6x} samples {sm:5d} executed {ex / chunks:6.1f}/chunk  {t[:90]}")
