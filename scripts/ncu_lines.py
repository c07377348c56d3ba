"""Per CUDA-C source line view of an ncu report (needs -lineinfo and --import-source on):
  python scripts/ncu_lines.py <report.ncu-rep> [N] [launch]
prints, per source file, the top-N lines by warp instructions executed with their share of stall samples."""
import csv
import subprocess
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
files, cur, hdr = {}, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        files.setdefault(cur, {})
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or cur is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    # rows with Address == '-' are the per-source-line aggregates
    if r[2] != "-":
        continue
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try:
        inst, smp = float(r[ii] or 0), float(r[si] or 0)
    except ValueError:
        continue
    d = files[cur].setdefault(int(r[0]), [r[1], 0.0, 0.0])
    d[1] += inst
    d[2] += smp
tot_i = sum(v[1] for f in files.values() for v in f.values())
tot_s = sum(v[2] for f in files.values() for v in f.values())
print(f"total warp instructions {tot_i:.3e}  samples {tot_s:.0f}")
for fname, lines in files.items():
    fi, fs = sum(v[1] for v in lines.values()), sum(v[2] for v in lines.values())
    if fi == 0:
        continue
    print(f"== {fname}: inst {100 * fi / tot_i:.1f}%  samples {100 * fs / max(tot_s, 1):.1f}%")
    for ln, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:topn]:
        if v[1] / tot_i < 0.002:
            break
        print(f"  {ln:4d} inst {100 * v[1] / tot_i:5.2f}%  smp {100 * v[2] / max(tot_s, 1):5.2f}%  {v[0].strip()[:110]}")
