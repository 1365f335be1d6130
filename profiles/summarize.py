#!/usr/bin/env python
"""Turn ncu artefacts brought back from the GPU box into the tracked summaries of this directory.

    python profiles/summarize.py launches <launch_list.csv> <out.md>      # per-launch device times of one step
    python profiles/summarize.py full <report.ncu-rep> <out.csv>           # key `--set full` metrics per kernel

The .ncu-rep files themselves stay in gpurun_out/ (scratch, tens of MB); what is committed here is derived
from them by this script on the build container (ncu -i ... --page raw --csv)."""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    seq = [(short(r[ki]), float(r[vi].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > vi]
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith("step_noise")]
    s0, s1 = (starts[-2], starts[-1]) if len(starts) >= 2 else (0, len(seq))
    step = seq[s0:s1]
    total = sum(v for _, v in step)
    agg = {}
    for n, v in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    with open(dst, "w") as f:
        f.write(f"One denoise step = {len(step)} launches, {total / 1e6:.3f} ms of device time under ncu "
                "(cold-cache, serialised: compare SHARES, not absolutes).\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {c} | {v / 1e3:.1f} | {100 * v / total:.1f} % |\n")
        f.write("\nLaunch order of the step:\n\n```\n")
        for n, v in step:
            f.write(f"{n:60s} {v / 1e3:10.1f} us\n")
        f.write("```\n")


def full(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([short(r[i]) if hdr[i] == "Kernel Name" else r[i] for i in idx])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
