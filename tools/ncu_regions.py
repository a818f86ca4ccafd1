"""Stall samples of one launch in an .ncu-rep aggregated by source-line ranges of the kernel file.
usage: ncu_regions.py report.ncu-rep launch_index file.cu:lo-hi[:name] ...   (ranges may repeat files)"""
import collections
import csv
import subprocess
import sys

rep, k = sys.argv[1], int(sys.argv[2])
specs = []
for a in sys.argv[3:]:
    parts = a.split(":")
    lo, hi = parts[1].split("-")
    specs.append((parts[0], int(lo), int(hi), parts[2] if len(parts) > 2 else a))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


cur, hdr, agg = None, None, {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    st = {h[6:]: num(r[i]) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h}
    agg[(cur, ln)] = (num(r[hdr["# Samples"]]), num(r[hdr["Instructions Executed"]]), st)
tot = sum(v[0] for v in agg.values()) or 1
seen = set()
print(f"launch {k}: {tot} samples")
for f, lo, hi, name in specs:
    keys = [key for key in agg if key[0] == f and lo <= key[1] <= hi]
    seen.update(keys)
    s = sum(agg[key][0] for key in keys)
    i = sum(agg[key][1] for key in keys)
    st = collections.Counter()
    for key in keys:
        st.update(agg[key][2])
    print(f"{name:26s} samples {s:5d} ({s / tot:.3f}) warp-inst {i:9d}  " + " ".join(f"{a}={b}" for a, b in st.most_common(6)))
rest = [key for key in agg if key not in seen]
print(f"{'(other)':26s} samples {sum(agg[key][0] for key in rest):5d}")
