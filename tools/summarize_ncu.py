#!/usr/bin/env python
"""Summarise an ncu report (or a launch-list csv) into a small markdown/CSV file for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep  > profiles/r01_kernels.md
    python tools/summarize_ncu.py --launches gpurun_out/launches.csv > profiles/r01_launches.md
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "  of which IMMA %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tensor-core smem operand pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU smem pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__sass_inst_executed_op_tmem_ldt.sum", "tcgen05.ld (LDTM) instructions"),
]


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print(f"# ncu --set full summary: `{path}`\n")
    print("Captured under the profiler (replayed, cold caches): use for counters and shares, not as bench values.\n")
    for r in rows[2:]:
        print(f"## {r[h.index('Kernel Name')]}\n")
        print("| metric | value |\n|---|---|")
        for m, label in METRICS:
            if m in h:
                i = h.index(m)
                print(f"| {label} (`{m}`) | {r[i]} {units[i]} |")
        print()


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        k = r[ki].split("(")[0]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list (`gpu__time_duration.sum`, --clock-control none): `{path}`\n")
    print("Per-launch times are cold-cache and serialised by the profiler: compare SHARES.\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in agg.items():
        print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / tot:.1%} |")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        kernels(sys.argv[1])
