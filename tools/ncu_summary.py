"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_bench_n1.md
    python tools/ncu_summary.py full gpurun_out/prof_head_fwd.ncu-rep profiles/r01_head_fwd_full.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^dram__bytes_(read|write)\.sum\.per_second$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"sm__pipe_tensor_cycles_active", r"sm__pipe_tensor_subpipe_hmma_cycles_active",
    r"^sm__inst_executed_pipe_tensor", r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^launch__(registers_per_thread|grid_size|block_size|"
    r"shared_mem_per_block_dynamic|occupancy_limit_shared_mem|waves_per_multiprocessor)$",
    r"^l1tex__m_xbar2l1tex_read_bytes\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?$",
    r"^l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sector_hit_rate\.pct$", r"^lts__t_bytes\.sum(\.per_second)?$",
    r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^gpc__cycles_elapsed\.max(\.per_second)?$",
    r"^sm__cycles_active\.avg$", r"^smsp__inst_executed\.sum$", r"^sm__inst_executed\.avg\.per_cycle_elapsed$",
    r"^smsp__warp_issue_stalled_.*_per_warp_active\.pct$", r"^l1tex__t_sector_hit_rate\.pct$",
    r"^sm__sass_inst_executed_op_(global|shared)_(ld|st|atom|red)\.sum$",
    r"^smsp__average_warps?_issue_stalled_.*\.ratio$", r"^smsp__average_warp_latency_issue_stalled",
]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}); per-launch times are cold-cache and serialised: compare SHARES\n\n")
        f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name[:90]}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / total:.1f}% |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src}\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')][:100]}  (id {r[0]})\n\n| metric | unit | value |\n|---|---|---:|\n")
            for i, h in enumerate(hdr):
                if any(re.search(p, h) for p in KEEP) and r[i] != "":
                    f.write(f"| {h} | {units[i]} | {r[i]} |\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
