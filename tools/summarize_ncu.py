#!/usr/bin/env python
"""Turn the ncu outputs of tools/ncu_capture.sh into the text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r01_g_launches.csv  > profiles/r01_g_launches_summary.txt
    python tools/summarize_ncu.py full     gpurun_out/r01_g_prof_mlp.ncu-rep > profiles/r01_g_mlp_full_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
           "lts__t_sectors_srcunit_tex_lookup_hit.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
           "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    k_name, k_val = head.index("Kernel Name"), head.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[k_name].split("(")[0]
        t = float(r[k_val].replace(",", "")) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    print(f"# launches {len(rows)}, total {total:.1f} us (cold-cache, serialised: compare SHARES)")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:80]:80s} n={n:4d} sum={t:10.1f}us avg={t / n:9.1f}us share={100 * t / total:5.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, rows = rows[0], rows[1], rows[2:]
    for r in rows:
        print("----")
        print(f"{'Kernel Name':70s} {r[head.index('Kernel Name')][:100]}")
        for m in METRICS:
            if m in head:
                i = head.index(m)
                print(f"{m:70s} {r[i]} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
