"""Summarise ncu reports / launch lists into small tracked files under profiles/.

    python tools/ncu_summary.py full   gpurun_out/prof.ncu-rep   profiles/r01_full_<tag>.md
    python tools/ncu_summary.py launch gpurun_out/launches.csv   profiles/r01_launches_<tag>.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep}`\n\n")
        f.write("Per captured launch (cold-cache, serialised replays: compare shares and ratios, not absolutes).\n"
                "`traffic` = dram__bytes_read.sum + dram__bytes_write.sum.\n\n")
        for r in rows[2:]:
            f.write(f"## {r[ki]}\n\n| metric | value | unit |\n|---|---|---|\n")
            vals = {}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    vals[m] = r[i]
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
            try:
                def gb(name):
                    i = hdr.index(name)
                    v = float(r[i])
                    u = units[i].lower()
                    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
                t = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
                f.write(f"| traffic (read+write) | {t / 1e9:.4f} | Gbyte |\n")
            except Exception:
                pass
            f.write("\n")


def launch(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui].strip(), 1e-3)   # -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    micro = {k: v for k, v in agg.items() if "k_fp64_" in k}      # bench.py's fp64 roof microbenchmark: not part of a step
    for k in micro:
        del agg[k]
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary of `{path}`\n\n`ncu --metrics gpu__time_duration.sum --clock-control none`; "
                f"first {sum(a[0] for a in agg.values())} launches of one bench step.  Times are cold-cache and serialised: "
                f"the SHARE of each kernel is what is comparable with bench.py's CUDA-event numbers.\n\n"
                f"| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |\n")
        if micro:
            f.write("\nNot part of a step (excluded from the shares): " +
                    ", ".join(f"`{k}` x{n} = {t:.0f} us" for k, (n, t) in micro.items()) +
                    " — `psfm_measure_dfma`, the fp64 roof microbenchmark `bench.py` runs once per process.\n")


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3])
