#!/usr/bin/env python
"""Turns the scratch artefacts of one gpurun call (gpurun_out/) into the small tracked summaries under profiles/.

    python profiles/summarise.py r01a            # tag = file-name prefix of this capture

Reads   gpurun_out/launches.csv                  (ncu --metrics gpu__time_duration.sum launch list of `bench.py`)
        gpurun_out/prof_*.ncu-rep                (ncu --set full captures)
        gpurun_out/bench.json, bench_ref.json, bench_extra.log
Writes  profiles/<tag>_launches.csv              kernel, launches, total / mean device time, share of all of OUR kernels
        profiles/<tag>_<capture>_raw.csv         the raw-page metrics the roofline uses, one row per profiled launch
        profiles/<tag>_bench.jsonl               the bench lines of the same call
        profiles/traffic.json                    dram bytes per launch of the dominant kernel (read by bench.py)
Needs ncu on PATH (it only reads reports; no GPU).
"""
import collections
import csv
import glob
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    return name.replace("evrep::", "").strip()


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"]) + " grid=" + r["Grid Size"].replace(" ", "")  # full-batch and e2e sub-batch launches differ in grid
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) / 1e3  # ns -> us
    ours = {k: v for k, v in agg.items() if k.startswith("k_")}
    tot = sum(v[1] for v in ours.values()) or 1.0
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write("kernel,launches,total_us,mean_us,share_of_our_kernels,grid,block\n")
        for k, v in agg.items():
            share = f"{v[1] / tot:.4f}" if k in ours else ""
            f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1] / v[0]:.2f},{share},\"{v[2]}\",\"{v[3]}\"\n")
    print("wrote", f"{tag}_launches.csv")


def raw(tag):
    traffic = {}
    tpath = os.path.join(PROF, "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        cap = os.path.basename(rep)[5:-8]
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        cols = [i for i, h in enumerate(hdr) if h in RAW_METRICS]
        kcol = hdr.index("Kernel Name")
        with open(os.path.join(PROF, f"{tag}_{cap}_raw.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["kernel"] + [f"{hdr[i]} [{units[i]}]" for i in cols])
            for row in rows[2:]:
                w.writerow([short(row[kcol])] + [row[i] for i in cols])
                d = dict(zip(hdr, row))
                u = dict(zip(hdr, units))

                def to_bytes(key):
                    v = float(d[key].replace(",", ""))
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]
                if "k_md_tile_static" in row[kcol] and cap.startswith("tile"):
                    traffic["ergo12_1mpx_b32"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
                    traffic["ergo12_1mpx_b32_source"] = f"profiles/{tag}_{cap}_raw.csv"
        print("wrote", f"{tag}_{cap}_raw.csv")
    json.dump(traffic, open(tpath, "w"), indent=1)


def bench(tag):
    with open(os.path.join(PROF, f"{tag}_bench.jsonl"), "w") as f:
        for name in ("bench.json", "bench_ref.json", "bench_extra.log", "bench_scale.log"):
            p = os.path.join(OUT, name)
            if os.path.exists(p):
                for line in open(p):
                    if line.startswith("{"):
                        f.write(line)
    print("wrote", f"{tag}_bench.jsonl")


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    raw(tag)
    bench(tag)
