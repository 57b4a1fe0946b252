#!/usr/bin/env python
"""
Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

    python tools/summarize_ncu.py r01 [traffic_key]   # expects gpurun_out/{launches,prof,bench}_r01.*
"""
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(REPO, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(REPO, "gpurun_out")

# ---- launch list ------------------------------------------------------------------------
lpath = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(lpath):
    rows = [r for r in csv.reader(open(lpath)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= col["Metric Value"]:
            continue
        name = r[col["Kernel Name"]].split("(")[0][:90]
        ns = float(r[col["Metric Value"]].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` over "
                "`python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline` (all launches of the "
                "process; cold-cache, serialised: compare shares).\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.2f} % |\n")
    print(open(os.path.join(out_dir, f"{tag}_launches.md")).read())

# ---- full captures of the dominant kernel of each workload ----------------------------------
WANT = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct",
    "sm__icc_request_hit_rate.pct",
]


def summarise(rep, out_md, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    with open(out_md, "w") as f:
        f.write(f"# ncu --set full: {title}\n\n| metric | unit | value |\n|---|---|---|\n")
        for k in WANT:
            if k in m:
                f.write(f"| {k} | {m[k][0]} | {m[k][1]} |\n")
        f.write("\n## warp stall reasons (warps stalled per issue-active cycle)\n\n| reason | value |\n|---|---:|\n")
        for h in hdr:
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                f.write(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} "
                        f"| {float(m[h][1]):.3f} |\n")

    def num(k):
        u, v = m[k]
        x = float(v.replace(",", ""))
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)

    fp64 = 0.0
    cyc = float(m["smsp__cycles_elapsed.avg"][1].replace(",", "")) if "smsp__cycles_elapsed.avg" in m else 0.0
    for op in ("dadd", "dmul", "dfma"):  # (--set full reports them per elapsed cycle, summed over the sub-partitions)
        k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"
        if k in m:
            fp64 += float(m[k][1].replace(",", "")) * cyc
    with open(out_md, "a") as f:
        f.write(f"\n## FP64\n\nfp64 thread-instructions (dadd + dmul + dfma, predicated on) per launch: {fp64:.6g}\n")
    grid = m.get("Grid Size", ("", ""))[1]
    return {"dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"), "fp64_inst": fp64,
            "kernel": m.get("Kernel Name", ("", ""))[1].split("(")[0], "grid": grid,
            "kernel_ms_under_ncu": float(m["gpu__time_duration.sum"][1].replace(",", "")) *
            {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(m["gpu__time_duration.sum"][0], 1.0)}


tpath = os.path.join(out_dir, "traffic.json")
tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
for w in ("c3", "c5", "c2", "c4"):
    rep = os.path.join(go, f"prof_{tag}_{w}.ncu-rep")
    if not os.path.exists(rep):
        continue
    traffic = summarise(rep, os.path.join(out_dir, f"{tag}_top_kernel_{w}.md"), f"dominant kernel of workload {w} ({tag})")
    # the bench looks the measurement up under "<workload>[_norss]_<N>x<M>x<T>" of its own config; the
    # captures of tools/profile_round2.sh run every workload at its benchmarked size
    bj = os.path.join(go, f"bench_{tag}.json")
    if os.path.exists(bj) and os.path.getsize(bj) > 2:
        line = json.loads(open(bj).read().strip().splitlines()[-1])
        cfg = line["config"] if w == "c3" else line.get("workloads", {}).get(w, {}).get("config")
        if cfg:
            key = f"{w}_{cfg['scenarios_per_gpu']}x{cfg['entities']}x{cfg['ticks']}"
            tj[key] = traffic
            print("per launch", traffic, "->", key)
    print(open(os.path.join(out_dir, f"{tag}_top_kernel_{w}.md")).read()[:1800])
json.dump(tj, open(tpath, "w"), indent=1, sort_keys=True)

bpath = os.path.join(go, f"bench_{tag}.json")
if os.path.exists(bpath):
    with open(os.path.join(out_dir, f"{tag}_bench.json"), "w") as f:
        f.write(open(bpath).read())
    rp = os.path.join(go, f"bench_ref_{tag}.json")
    if os.path.exists(rp):
        with open(os.path.join(out_dir, f"{tag}_bench_reference_arm.json"), "w") as f:
            f.write(open(rp).read())
    for w in ("c2", "c4", "c5", "c3_norss"):
        bp = os.path.join(go, f"bench_{tag}_{w}.json")
        if os.path.exists(bp):
            with open(os.path.join(out_dir, f"{tag}_bench_{w}.json"), "w") as f:
                f.write(open(bp).read())
