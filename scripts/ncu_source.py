"""Hot-spot view of an ncu source page: python scripts/ncu_source.py <report.ncu-rep> [N] [launch]
prints the opcode mix with instruction / stall-sample shares, the per-phase shares of the two-phase conv kernels (regions
split at the first BAR.SYNC and the first / last dense FFMA2 block) and the top-N SASS lines.  `launch` selects one launch of
a multi-launch report (default 0); ncu's CSV export lists a kernel's SASS twice, the duplicate half is dropped."""
import csv
import subprocess
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) > idx["Instructions Executed"] and r[0].startswith("0x") or (r and r[0].isdigit())]
half = len(data) // 2
if half > 0 and [r[idx["Source"]] for r in data[:half]] == [r[idx["Source"]] for r in data[half:2 * half]]:
    data = data[:half]


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

# per-phase shares of the two-phase conv kernels
src = [r[idx["Source"]] for r in data]
bars = [i for i, t in enumerate(src) if "BAR.SYNC" in t]
ff = [i for i, t in enumerate(src) if "FFMA2" in t]
if bars and ff:
    b0, l0 = bars[0], ff[0]
    end = l0
    for i in range(l0, max(l0, len(data) - 50), 10):
        if sum("FFMA2" in t for t in src[i:i + 50]) >= 15:
            end = i + 50
    print("--- regions")
    for name, (a, b) in (("phase 1 (before the first BAR.SYNC)", (0, b0)), ("barrier + phase-2 prologue", (b0, l0)),
                         ("phase 2 main loop (dense FFMA2)", (l0, end)), ("epilogue", (end, len(data)))):
        ins = sum(f(r, "Instructions Executed") for r in data[a:b])
        smp = sum(f(r, "# Samples") for r in data[a:b])
        print(f"  {name:40s} inst {100*ins/tot_i:5.1f} %  samples {100*smp/max(tot_s,1):5.1f} %")
