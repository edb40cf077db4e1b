#!/usr/bin/env python
"""Condenses an Nsight Compute report (.ncu-rep, `ncu --set full`) or a launch list
(`--metrics gpu__time_duration.sum --csv`) into the small text tables kept under profiles/.

  python tools/ncu_summary.py report  gpurun_out/x.ncu-rep   > profiles/x.md
  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/launches.md
  python tools/ncu_summary.py traffic  gpurun_out/x.ncu-rep MESH ORDERING > profiles/ncu_traffic.json
      (dram__bytes_read.sum + dram__bytes_write.sum per launch of the hot kernels: what bench.py reports as roofline.traffic)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % (active)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles %"),
    ("smsp__inst_executed_pipe_fp64.sum", "FP64 warp insts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__shared_mem_per_block_static", "static smem"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar / issue"),
]


def short(name):
    name = name.replace("void ", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "").replace("nsx::", "")
    return re.sub(r"\(.*", "", name)


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(head)}
    print(f"# {path}\n")
    print("`ncu --set full --clock-control none` capture; per-launch values (cold caches, serialised replay).\n")
    cols = [(k, lab) for k, lab in KEYS if k in idx]
    print("| kernel | " + " | ".join(f"{lab} [{units[idx[k]]}]" if units[idx[k]] else lab for k, lab in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in body:
        print("| " + short(r[idx["Kernel Name"]]) + " | " + " | ".join(r[idx[k]] for k, _ in cols) + " |")


def launches(path):
    lines = open(path).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[i:]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        a = agg[short(r["Kernel Name"])]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}\n")
    print(f"`ncu --metrics gpu__time_duration.sum --clock-control none`: {sum(v[0] for v in agg.values())} launches, "
          f"{tot / 1e6:.3f} ms of kernel time (per-launch times are cold-cache and serialised: read the SHARES).\n")
    print("| kernel | launches | total ms | share % | avg us |")
    print("|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.2f} | {v[1] / v[0] / 1e3:.1f} |")


def traffic(path, mesh, ordering):
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = {n: i for i, n in enumerate(head)}

    def to_bytes(v, unit):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

    per = collections.defaultdict(list)
    for r in body:
        b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        per[short(r[idx["Kernel Name"]])].append(b)
    # k_spmv_tma<1, 0> = general rows (the Jacobian block product; also the F product when no node view is active),
    # k_spmv_tma<1, 1|2> = node rows (the F product of the inner solves in the Stokes-type branches)
    node = [v for k, vs in per.items() if k.startswith("k_spmv_tma<1, 1") or k.startswith("k_spmv_tma<1, 2") for v in vs]
    general = [v for k, vs in per.items() if k.startswith("k_spmv_tma<1, 0") for v in vs]
    sweep = [v for k, vs in per.items() if k.startswith("k_sweep_block") for v in vs]
    med = lambda v: sorted(v)[len(v) // 2]
    res = {}
    if sweep:
        res["sgs_F"] = med(sweep)
    if node or general:
        res["spmv_F"] = med(node) if node else med(general)
    if general and node:
        res["block_spmv"] = max(general)
    print(json.dumps({"mesh": mesh, "ordering": int(ordering), "source": path.split("/")[-1], "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch (median over the captured launches; block_spmv = the largest launch on general rows, when one was captured)",
                      "bytes_per_launch": res, "all_kernels_median_bytes": {k: sorted(v)[len(v) // 2] for k, v in per.items()}}, indent=1))


if __name__ == "__main__":
    {"report": report, "launches": launches, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
