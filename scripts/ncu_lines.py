"""Joins an ncu SASS source page with nvdisasm line info: samples / instructions per CUDA source line.
  python scripts/ncu_lines.py <report.ncu-rep> <mangled kernel substring> [topN]
Needs the .so the report was taken from (dmcf_b200/lib/libdmcf_b200.so, built with -lineinfo)."""
import csv, glob, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "dmcf_b200/lib/libdmcf_b200.so")], cwd=tmp, capture_output=True)
addr2line = {}
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if kern not in dis:
        continue
    in_fn, cur = False, None
    for line in dis.splitlines():
        if line.startswith(".text.") and line.endswith(":"):
            in_fn = kern in line
            continue
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
        if m and cur:
            addr2line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
idx = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
base = min(int(r[0], 16) for r in data if r and r[0].startswith("0x")) if data and data[0][0].startswith("0x") else 0
agg = {}
ts = ti = 0.0
for n, r in enumerate(data):
    try:
        s = float(r[idx["# Samples"]] or 0); i = float(r[idx["Instructions Executed"]] or 0)
    except (ValueError, IndexError):
        continue
    off = (int(r[0], 16) - base) if r[0].startswith("0x") else n * 16
    key = addr2line.get(off, ("?", 0))
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += s; a[1] += i; ts += s; ti += i
srcs = {}
print(f"samples {ts:.0f}  warp-instructions {ti:.3e}")
for (f, ln), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    if f not in srcs:
        p = os.path.join(root, "dmcf_b200/csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][ln - 1].strip()[:90] if 0 < ln <= len(srcs[f]) else ""
    print(f"{100*s/ts:5.1f}% smp {100*i/ti:5.1f}% inst  {f}:{ln:<4} {text}")
