#!/usr/bin/env python
"""Reduce the ncu outputs of tools/gpu_ncu_r2.sh to the tracked summaries under profiles/.
usage: python tools/summarize_ncu_r2.py <tag>   (reads gpurun_out/{launches,launches128,prof,prof128,profc}_<tag>.*)"""
import csv, io, json, re, subprocess, sys
from collections import OrderedDict
tag = sys.argv[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__shared_mem_per_block_dynamic',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']
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
    return agg, tot


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
    return caps


def dram(c):
    return sum(float(c[k]["value"].replace(',', '')) * UNIT[c[k]["unit"]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))


base = "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --lean --nprobe 8 --profile-window"
a, t = launches(f'gpurun_out/launches_{tag}.csv', 'profiles/r2_launches.json', "ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none " + base)
print("10k step", round(t, 1), "us", {k: (v["launches"], v["us"], v["share"]) for k, v in a.items()})
a, t = launches(f'gpurun_out/launches128_{tag}.csv', 'profiles/r2_launches_nq128.json', "ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none " + base + " --profile-nq 128")
print("128 step", round(t, 1), "us", {k: (v["launches"], v["us"], v["share"]) for k, v in a.items()})
full = "ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_tc_kernel -c 2 "
c10 = captures(f'gpurun_out/prof_{tag}.ncu-rep', 'profiles/r2_ncu_scan_tc.json', full + base, ["bounds launch (minima of the heads of each query's four nearest lists)", "main launch"])
c128 = captures(f'gpurun_out/prof128_{tag}.ncu-rep', 'profiles/r2_ncu_scan_tc_nq128.json', full + base + " --profile-nq 128",
                ["bounds launch (128 queries: 512 tiles per query)", "main launch (CTA-local top-k sets)"])
import os
if os.path.exists(f'gpurun_out/profc_{tag}.ncu-rep'):  # (the closing visit, tools/gpu_final2.sh, does not repeat the coarse capture)
  captures(f'gpurun_out/profc_{tag}.ncu-rep', 'profiles/r2_ncu_coarse_c5shape.json',
         "ncu --profile-from-start off --set full --clock-control none -k regex:coarse_dist_kernel|scan_tc_kernel|select_topk_kernel|select_small_kernel|finalize_kernel -c 8 python tools/coarse_ncu.py",
         ["exact: coarse_dist_kernel", "exact: select_topk_kernel", "filter: scan_tc_kernel bounds pass over the centroid table", "filter: select_small_kernel",
          "filter: scan_tc_kernel frozen pass", "filter: finalize_kernel (exact re-check + probe order)"])
w = {"n": 1000000, "d": 128, "nq": 10000, "k": 10, "nlist": 1024, "seed": 42, "nprobe": 8}
json.dump({"workload": w, "kernel": "scan_tc_kernel, bounds + main launch of one search step",
           "dram_bytes": sum(dram(c) for c in c10), "dram_bytes_nq128": sum(dram(c) for c in c128),
           "source": "profiles/r2_ncu_scan_tc.json and profiles/r2_ncu_scan_tc_nq128.json (dram__bytes_read.sum + dram__bytes_write.sum of both launches)"},
          open('profiles/ncu_traffic.json', 'w'), indent=1)
print(open('profiles/ncu_traffic.json').read())
