#!/usr/bin/env python
"""Reduce the ncu outputs of tools/gpu_final3.sh to the tracked summaries under profiles/ (r3_*).
usage: python tools/summarize_ncu_r3.py <tag>"""
import csv, io, json, os, re, subprocess, sys
from collections import OrderedDict
tag = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic']
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}


def launches(path, out, cmd):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]; ik, iv, ig, ib = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size')
    ls = [{"kernel": r[ik], "grid": r[ig], "block": r[ib], "us": float(r[iv].replace(',', '')) / 1e3} for r in rows[1:]]
    tot = sum(l["us"] for l in ls)
    agg = OrderedDict()
    for l in ls:
        n = re.sub(r'\(.*', '', l["kernel"]).replace('void ', '')
        a = agg.setdefault(n, {"launches": 0, "us": 0.0}); a["launches"] += 1; a["us"] += l["us"]
    for a in agg.values(): a["share"] = round(a["us"] / tot, 4); a["us"] = round(a["us"], 1)
    json.dump({"command": cmd, "note": "one warmed-up search step (cudaProfilerStart/Stop window); per-launch times are serialised and cold-cache: compare shares",
               "step_us": round(tot, 1), "n_launches": len(ls), "by_kernel": agg, "launches": ls}, open(out, 'w'), indent=1)
    print(out, round(tot, 1), "us", len(ls), "launches", {k: (v["launches"], v["us"], v["share"]) for k, v in agg.items() if v["share"] > 0.01})


def captures(rep, out, cmd, what):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw))); hdr, units = r[0], r[1]
    caps = []
    for row in r[2:]:
        d = {"kernel": row[hdr.index('Kernel Name')]}
        for w in WANT:
            if w in hdr: d[w] = {"value": row[hdr.index(w)], "unit": units[hdr.index(w)]}
        caps.append(d)
    json.dump({"command": cmd, "launches": what, "captures": caps}, open(out, 'w'), indent=1)
    for c in caps:
        print(out, c["kernel"][:60], {k.split('.')[0]: v["value"] + ' ' + v["unit"] for k, v in c.items() if k != "kernel" and k in WANT[:4]})
    return caps


base = "python bench.py --steps 1 --warmup 4 --no-cpu-baseline --lean --nprobe 8 --profile-window"
ll = "ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none "
if os.path.exists(f'gpurun_out/launches_graph_{tag}.csv'):
    launches(f'gpurun_out/launches_graph_{tag}.csv', 'profiles/r3_launches_graph.json', ll + base + "   (graph replay on: the kernels are nodes of one cudaGraphLaunch)")
if os.path.exists(f'gpurun_out/launches_{tag}.csv'):
    launches(f'gpurun_out/launches_{tag}.csv', 'profiles/r3_launches.json', "VIDX_GRAPH=0 " + ll + base)
full = "VIDX_GRAPH=0 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 2 -c 2 "
if os.path.exists(f'gpurun_out/prof_{tag}.ncu-rep'):
    captures(f'gpurun_out/prof_{tag}.ncu-rep', 'profiles/r3_ncu_scan_tc.json', full + base,
             ["bounds launch (minima of the heads of each query's four nearest lists)", "main launch"])
if os.path.exists(f'gpurun_out/prof768_{tag}.ncu-rep'):
    captures(f'gpurun_out/prof768_{tag}.ncu-rep', 'profiles/r3_ncu_scan_tc_d768.json',
             "VIDX_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel|tc_atile_kernel -c 7 "
             "python tools/bench_yaml_grid.py --dims 768 --counts 100000 --nprobes 8 --reps 1",
             "first search of the tool (100 000 x 768, nlist 1260, 10 000 queries, n_probe 8), in launch order: coarse filter (tc_atile, bounds pass, "
             "frozen pass over the centroid table, streamed query tiles), then the list scan (tc_atile x 2, seeding bounds launch, main launch)")
