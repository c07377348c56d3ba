"""Hot-spot view of an ncu source page: python scripts/ncu_source.py <report.ncu-rep> [N]
prints instruction-count / stall-sample shares grouped by SASS region, and the top-N SASS lines."""
import csv
import subprocess
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) > idx["Instructions Executed"] and r[0].startswith("0x") or (r and r[0].isdigit())]
def f(r, k):
    try:
        return float(r[idx[k]] or 0)
    except (ValueError, IndexError):
        return 0.0
tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_s = sum(f(r, "# Samples") for r in data)
print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}, SASS lines {len(data)}")
# opcode mix
mix = {}
for r in data:
    op = r[idx["Source"]].strip().split()[0] if r[idx["Source"]].strip() else "?"
    if op.startswith("@"):
        op = r[idx["Source"]].strip().split()[1]
    op = op.split(".")[0]
    m = mix.setdefault(op, [0.0, 0.0, 0.0])
    m[0] += f(r, "Instructions Executed"); m[1] += f(r, "# Samples"); m[2] += f(r, "L1 Wavefronts Shared")
print("opcode  inst%  samples%  smem_wavefronts")
for op, m in sorted(mix.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{op:10s} {100*m[0]/tot_i:6.1f} {100*m[1]/max(tot_s,1):6.1f} {m[2]:.3e}")
print("--- top lines by samples")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:topn]:
    stalls = {k: f(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*f(r,'# Samples')/max(tot_s,1):5.1f}% inst {100*f(r,'Instructions Executed')/tot_i:4.1f}%  {r[idx['Source']][:70]:70s} {top}")
