"""Per-source-line view of one launch in an .ncu-rep: stall samples, warp instructions, shared-memory
wavefront excess.  usage: ncu_lines.py report.ncu-rep [launch_index] [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, collections.Counter()])   # samples, inst, wavefronts, ideal, stalls
text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        stall_cols = [(h, i) for i, h in enumerate(r) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    text[key] = r[1].strip()[:90]
    a = agg[key]
    def num(name):
        try:
            return int(r[hdr[name]])
        except (ValueError, KeyError, IndexError):
            return 0
    a[0] += num("# Samples")
    a[1] += num("Instructions Executed")
    a[2] += num("L1 Wavefronts Shared")
    a[3] += num("L1 Wavefronts Shared Ideal")
    for h, i in stall_cols:
        try:
            a[4][h] += int(r[i])
        except (ValueError, IndexError):
            pass
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"launch {k}: {tot_s} samples, {tot_i} warp instructions")
print("share_samples share_inst  smem_wavefronts/ideal  file:line  top stalls | source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join(f"{h[6:]}={v}" for h, v in a[4].most_common(3))
    print(f"{a[0]/tot_s:6.3f} {a[1]/tot_i:6.3f} {a[2]:>9d}/{a[3]:<9d} {key[0]}:{key[1]:<4d} {st:40s} | {text[key]}")
