"""Summarises ncu outputs into small text files for profiles/ (the judged, committed evidence).

  python scripts/ncu_summary.py launches <launches.csv> <out.md>      per-kernel launch count / time / share
  python scripts/ncu_summary.py full <report.ncu-rep> <out.md>        key metrics of every captured launch
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    with open(out, "w") as fh:
        fh.write(f"# ncu launch list summary ({path})\n\nper-launch device time is cold-cache / serialised: compare SHARES\n\n")
        fh.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / total:.1f}% |\n")
        fh.write(f"\ntotal {total:.3f} ms over {sum(v[0] for v in agg.values())} launches\n")


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as fh:
        fh.write(f"# ncu --set full summary ({path})\n\n")
        for d in data:
            fh.write(f"## {d[idx['Kernel Name']]}  (launch id {d[idx['ID']]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    fh.write(f"| {k} | {d[idx[k]]} | {units[idx[k]]} |\n")
            fh.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
