#!/usr/bin/env python
"""Turn the ncu artefacts of a round into the committed text summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_rNN.csv  > profiles/rNN_launches.md
    python profiles/summarize.py full     gpurun_out/prof_rNN.ncu-rep  > profiles/rNN_full.md

`launches` = the `--metrics gpu__time_duration.sum --clock-control none` pass over one bench.py run
(per-launch times are cold-cache and serialised: compare SHARES, not absolutes).
`full` = one `--set full` capture; needs `ncu` on PATH to read the report.
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | avg us | share | grid (last) | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f | %.1f %% | %s | %s |" % (k, a[0], a[1] / 1e6, a[1] / a[0] / 1e3,
                                                                   100 * a[1] / tot, a[2], a[3]))
    print("\ntotal device time of the listed launches: %.3f ms" % (tot / 1e6))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    seen = set()
    for r in rows[2:]:
        name = r[ni].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print("### `%s`\n" % name)
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("| %s | %s | %s |" % (k, r[i], units[i]))
        print()


def traffic(path, workload="cfg3"):
    """Single-pass DRAM capture (dram__bytes_read/write + duration per launch) -> the JSON bench.py reads."""
    import json
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    fam_of = [("k_inspectors", "inspector"), ("k_chan_ifft", "chan_ifft"), ("k_cols256", "fft_cols"),
              ("k_rows256<1>", "fft_rows_chan"), ("k_rows256<(int)1>", "fft_rows_chan"),
              ("k_rows256<0>", "fft_rows_psd"), ("k_rows256<(int)0>", "fft_rows_psd"), ("k_sym_pack", "sym_pack")]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    per = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        fam = next((f for k, f in fam_of if k in r[ki]), None)
        if fam is None:
            continue
        a = per.setdefault(fam, {}).setdefault(r[ii], {})
        a[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    out = {"workload": workload, "source": path + ": ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
           "gpu__time_duration.sum --clock-control none --cache-control none (single pass, no replay), "
           "bench.py --workload %s --steps 2 --warmup 3" % workload, "kernels": {}}
    for fam, ls in per.items():
        n = len(ls)
        rd = sum(l.get("dram__bytes_read.sum", 0) for l in ls.values()) / n
        wr = sum(l.get("dram__bytes_write.sum", 0) for l in ls.values()) / n
        ms = sum(l.get("gpu__time_duration.sum", 0) for l in ls.values()) / n
        out["kernels"][fam] = {"dram_bytes_per_launch": round(rd + wr), "read": round(rd), "write": round(wr),
                               "launch_ms_under_ncu": round(ms, 4), "launches": n}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
