#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small tracked summaries in profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  STEPS  > profiles/rNN_launches.md
  python profiles/summarize.py full     gpurun_out/prof.ncu-rep         > profiles/rNN_ncu_full_<kernel>.csv

`launches` = the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised: read SHARES).
`full`     = one `--set full` capture; keeps the metrics the roofline argument needs.
`traffic`  = profiles/ncu_traffic.json (DRAM bytes per launch of each captured kernel, stamped with the source hash):
  python profiles/summarize.py traffic profiles/ncu_traffic.json gpurun_out/a.ncu-rep gpurun_out/b.ncu-rep=cfg2 ...
"""
import collections
import csv
import subprocess
import sys

FULL_KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_barrier", "sass__inst_executed_local_loads",
    "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex.max.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]


def launches(path, steps):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"| us/step | launches/step | share | kernel |\n|---:|---:|---:|---|")
    for name, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {t / steps:.1f} | {c / steps:.1f} | {100 * t / tot:.1f}% | `{name}` |")
    print(f"| {tot / steps:.1f} | {sum(a[0] for a in agg.values()) / steps:.1f} | 100% | total |")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(1, len(rows) - 1)])
    for k in FULL_KEEP:
        if k in H:
            i = H.index(k)
            w.writerow([k, rows[1][i]] + [r[i] for r in rows[2:]])


def short_name(k):
    """`void fr::seg_chunk_kernel<fr::UserPol<1, 3>>(...)` -> `seg_chunk_kernel<UserPol>`; other kernels: bare name."""
    import re
    k = k.replace("void ", "").replace("fr::", "").split("(")[0]
    m = re.match(r"(seg_\w+)<(\w+)", k)
    return f"{m.group(1)}<{m.group(2)}>" if m else k.split("<")[0]


def traffic(out_path, reps):
    """profiles/ncu_traffic.json from `ncu --set full` captures: per kernel (first launch in each capture) the DRAM
    bytes read + written, stamped with the hash of the sources of the CURRENT tree (bench.py only uses the file while
    that hash matches the library it runs).  A capture argument may be `file.ncu-rep=suffix`: the suffix is appended to
    the kernel name (`catalog_gemm_kernel:cfg2`)."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from foodrec_b200 import _build
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    kernels = {}
    for rep in reps:
        rep, _, suffix = rep.partition("=")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        H, U = rows[0], rows[1]
        ir, iw, ik = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("Kernel Name")
        it = H.index("gpu__time_duration.sum")
        for r in rows[2:]:
            name = short_name(r[ik]) + (":" + suffix if suffix else "")
            if name in kernels:
                continue
            b = float(r[ir].replace(",", "")) * unit[U[ir]] + float(r[iw].replace(",", "")) * unit[U[iw]]
            kernels[name] = {"dram_bytes": b, "dram_read": float(r[ir].replace(",", "")) * unit[U[ir]],
                             "duration": r[it] + " " + U[it], "capture": os.path.basename(rep)}
    json.dump({"build": _build.source_hash(), "kernels": kernels}, open(out_path, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], float(sys.argv[3]))
    elif sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3:])
    else:
        full(sys.argv[2])
