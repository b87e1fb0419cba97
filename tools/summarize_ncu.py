#!/usr/bin/env python
"""Turn the ncu outputs of tools/gpu_round.sh into the tracked summaries under profiles/.
usage: python tools/summarize_ncu.py <tag> <out-prefix>   (reads gpurun_out/{prof,launches,bench}_<tag>.*)"""
import csv, json, subprocess, sys, io, re
from collections import OrderedDict
tag, out = sys.argv[1], sys.argv[2]
# 1. launch list of one warmed-up step
rows = [r for r in csv.reader(open(f'gpurun_out/launches_{tag}.csv')) if len(r) > 5]
h = rows[0]; ik, iv, ig, ib = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size')
launches = [{"kernel": r[ik], "grid": r[ig], "block": r[ib], "us": float(r[iv].replace(',', '')) / 1e3} for r in rows[1:]]
tot = sum(l["us"] for l in launches)
agg = OrderedDict()
for l in launches:
    n = re.sub(r'\(.*', '', l["kernel"]).replace('void ', '')
    a = agg.setdefault(n, {"launches": 0, "us": 0.0}); a["launches"] += 1; a["us"] += l["us"]
for a in agg.values(): a["share"] = round(a["us"] / tot, 4); a["us"] = round(a["us"], 1)
json.dump({"command": "ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline --nprobe 8 --profile-window",
           "note": "one warmed-up search step (cudaProfilerStart/Stop window); per-launch times are serialised and cold-cache: compare shares",
           "step_us": round(tot, 1), "by_kernel": agg, "launches": launches}, open(f'{out}_launches.json', 'w'), indent=1)
# 2. full capture of the scan kernel (two launches: seeding pass, main pass)
raw = subprocess.run(['ncu', '-i', f'gpurun_out/prof_{tag}.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); hdr, units = r[0], r[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
caps = []
for row in r[2:]:
    d = {"kernel": row[hdr.index('Kernel Name')]}
    for w in want:
        if w in hdr: d[w] = {"value": row[hdr.index(w)], "unit": units[hdr.index(w)]}
    caps.append(d)
json.dump({"command": "ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:scan_tc_kernel -c 2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --nprobe 8 --profile-window",
           "launches": ["seeding bounds pass (minima only, heads of each query's four nearest lists)", "main pass"], "captures": caps}, open(f'{out}_ncu_scan_tc.json', 'w'), indent=1)
# 3. dram traffic per step for bench.py's roofline.traffic
def mb(c, k): return float(c[k]["value"].replace(',', '')) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[c[k]["unit"]]
bench = [json.loads(l) for l in open(f'gpurun_out/bench_{tag}.json') if l.startswith('{')][-1]
json.dump({"workload": {k: bench["config"][k] for k in ("n", "d", "nq", "k", "nlist", "seed", "nprobe")},
           "kernel": "scan_tc_kernel, seeding + main launch of one search step",
           "dram_bytes": sum(mb(c, 'dram__bytes_read.sum') + mb(c, 'dram__bytes_write.sum') for c in caps),
           "source": f"{out}_ncu_scan_tc.json (dram__bytes_read.sum + dram__bytes_write.sum)"}, open('profiles/ncu_traffic.json', 'w'), indent=1)
json.dump(bench, open(f'{out}_bench.json', 'w'), indent=1)
print(open('profiles/ncu_traffic.json').read()); print({k: v for k, v in agg.items()})
