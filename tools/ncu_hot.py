"""Print the hottest SASS lines (warp-stall samples) of every kernel in an ncu report.
usage: python tools/ncu_hot.py report.ncu-rep [min_share]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.015
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("==", {k: r[h.index(k)] for k in want if k in h})
    st = {n.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i]) for i, n in enumerate(h)
          if n.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in n and r[i]}
    tot = sum(st.values()) or 1
    print("   stalls:", {k: round(v / tot, 3) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; block = []
def flush():
    if not block: return
    hi = block[0]; col = hi.index("Warp Stall Sampling (All Samples)"); s = hi.index("Source")
    data = []
    for idx, x in enumerate(block[1:]):
        try: data.append((int(x[col]), idx, x[s]))
        except Exception: pass
    tot = sum(a for a, _, _ in data) or 1
    print("--", cur, "samples", tot)
    for a, idx, t in data:
        if a > tot * thr: print(f"   {idx:5d} {a:6d} {100*a/tot:5.1f}%  {t[:110]}")
for r in rows:
    if r and r[0] == "Kernel Name":
        flush(); cur = r[1][:80]; block = []
    elif "Source" in r and "Address" in r:
        block = [r]
    elif block:
        block.append(r)
flush()
