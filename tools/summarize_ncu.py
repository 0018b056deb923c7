"""Summarise ncu outputs for profiles/ (run here, no GPU needed).
    python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.md
    python tools/summarize_ncu.py full gpurun_out/prof_corr.ncu-rep > profiles/r1_ncu_corr.md"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    unit = [r for r in rows[1:] if r[im] == "gpu__time_duration.sum"][0][hdr.index("Metric Unit")]
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total %s | share |" % unit)
    print("|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))
    print("\ntotal %.1f %s over %d launches (cold-cache, serialised: compare shares, not absolutes)" % (tot, unit, sum(v[0] for v in agg.values())))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("### %s" % r[hdr.index("Kernel Name")][:160])
        print("| metric | value | unit |\n|---|---|---|")
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                print("| %s | %s | %s |" % (h, v, u))
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
