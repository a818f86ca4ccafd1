"""Summarise an .ncu-rep (raw page + SASS hot spots) into text for profiles/."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
idx = {h: i for i, h in enumerate(H)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in idx:
            print(f"{w:75s} {r[idx[w]]:>18s} {rows[1][idx[w]]}")
    st = [(h, r[idx[h]]) for h in H if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    st = sorted(st, key=lambda x: -float(x[1].replace(",", "") or 0))[:8]
    print("top stalls (warps per issue):", ", ".join(f"{h.split('stalled_')[1].replace('_per_issue_active.ratio','')}={v}" for h, v in st))
if len(sys.argv) > 2:
    k = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    SH = srows[1]
    I = {h: i for i, h in enumerate(SH)}
    data = srows[2:]
    tot_s = sum(int(r[I["# Samples"]]) for r in data) or 1
    tot_i = sum(int(r[I["Instructions Executed"]]) for r in data) or 1
    print("=" * 100)
    print(f"SASS profile of launch {k}: {len(data)} instructions, {tot_i} warp-instructions executed, {tot_s} samples")
    op_i, op_s = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[I["Source"]].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        op_i[op] += int(r[I["Instructions Executed"]])
        op_s[op] += int(r[I["# Samples"]])
    for k2, v in op_i.most_common(18):
        print(f"  {k2:12s} inst {v / tot_i:6.3f}  samples {op_s[k2] / tot_s:6.3f}")
    print("hottest SASS by stall samples:")
    for r in sorted(data, key=lambda r: -int(r[I["# Samples"]]))[:25]:
        print(f"  {int(r[I['# Samples']]) / tot_s:6.3f}  {r[I['Source']].strip()[:90]}")
