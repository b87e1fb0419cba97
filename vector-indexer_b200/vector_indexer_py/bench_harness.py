"""The reference's benchmark harness, run against the B200 library.

bench/faiss_bench_official/bench_all_ivf.py measures an index the Faiss way (its eval_setting, :283-363): search the whole
query batch over and over until a minimum wall time has passed, report ms/query and QPS from the mean, and R@1 / R@10 /
R@100 = the fraction of queries whose TRUE nearest neighbour shows up among the first 1 / 10 / 100 results; one row per
n_probe of a sweep (:427-480); results go to `faiss_bench_results.json` and a markdown table (:514-533).

This module does the same sweep through the same adapter surface (`index.nprobe = p; D, I = index.search(xq, k)`), writes
result files with the same keys and the same table columns, and takes its ground truth from an exact float64 brute force
(torch when a GPU is there, numpy otherwise) instead of faiss.IndexFlatL2, which is not installed here.  Data: the
harness's synthetic generator (:67-69) -- one `default_rng(seed)` stream, xb first, then xq.

    python -m vector_indexer_py.bench_harness --n 100000 --d 128 --nq 1000 --k 100 --output-dir out/
Defaults are the reference's (scripts/run_faiss_bench.sh:51-57): n 100 000, d 128, nq 1000, k 100,
n_probe 1..64, 3 s per setting, seed 42.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

from . import build, suggest_nlist
from .faiss_adapter import VectorIndexerFaissAdapter, recall_at_ranks

RANKS = (1, 10, 100)


def synthetic_dataset(n, d, nq, seed=42):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((n, d)).astype(np.float32)
    xq = rng.standard_normal((nq, d)).astype(np.float32)
    return xb, xq


def exact_ground_truth(xb, xq, k, block=65536):
    """Ids of the k exact nearest neighbours (float64 squared L2), blockwise; what IndexFlatL2 provides in the reference."""
    try:
        import torch
        if torch.cuda.is_available():
            dev = torch.device("cuda")
            q = torch.from_numpy(xq).to(dev).double()
            qn = (q * q).sum(1, keepdim=True)
            best_d = torch.full((len(xq), k), float("inf"), dtype=torch.float64, device=dev)
            best_i = torch.full((len(xq), k), -1, dtype=torch.int64, device=dev)
            for b0 in range(0, len(xb), block):
                x = torch.from_numpy(xb[b0:b0 + block]).to(dev).double()
                dist = qn - 2.0 * (q @ x.T) + (x * x).sum(1)[None, :]
                dv, di = torch.topk(dist, min(k, len(x)), dim=1, largest=False)
                cd, ci = torch.cat([best_d, dv], 1), torch.cat([best_i, di + b0], 1)
                o = torch.argsort(cd, dim=1, stable=True)[:, :k]
                best_d, best_i = torch.gather(cd, 1, o), torch.gather(ci, 1, o)
            return best_i.cpu().numpy()
    except ImportError:
        pass
    q = xq.astype(np.float64)
    best_d = np.full((len(xq), k), np.inf)
    best_i = np.full((len(xq), k), -1, np.int64)
    for b0 in range(0, len(xb), block):
        x = xb[b0:b0 + block].astype(np.float64)
        dist = (q * q).sum(1)[:, None] - 2.0 * q @ x.T + (x * x).sum(1)[None, :]
        cd = np.concatenate([best_d, dist], 1)
        ci = np.concatenate([best_i, np.arange(b0, b0 + len(x))[None, :].repeat(len(xq), 0)], 1)
        o = np.argsort(cd, 1, kind="stable")[:, :k]
        best_d, best_i = np.take_along_axis(cd, o, 1), np.take_along_axis(ci, o, 1)
    return best_i


def eval_setting(index, xq, gt, k, min_time, out=sys.stdout):
    """One row of the sweep: repeat the batch until `min_time` seconds are over (at least once)."""
    nq = len(xq)
    nrun, t0 = 0, time.time()
    while True:
        _, I = index.search(xq, k)
        nrun += 1
        t1 = time.time()
        if t1 - t0 > min_time:
            break
    ms_per_query = (t1 - t0) * 1000.0 / nq / nrun
    res = {"ms_per_query": ms_per_query, "qps": 1000.0 / ms_per_query, "nrun": nrun,
           "recalls": recall_at_ranks(I, gt, ranks=RANKS)}
    cells = "  ".join("R@%-3d=%.4f" % (r, v) for r, v in res["recalls"].items())
    print("%s    %9.3f ms/q  %9.1f QPS  (nrun=%d)" % (cells, ms_per_query, res["qps"], nrun), file=out)
    return res


def benchmark_vector_indexer(xb, xq, gt, k, nprobes, min_time, work_dir=None, out=sys.stdout, **build_kw):
    n, d = xb.shape
    nlist = build_kw.get("nlist") or suggest_nlist(n)
    print(f"vector_indexer on B200: n={n}, d={d}, nlist={nlist}", file=out)
    t0 = time.time()
    idx = build(xb, work_dir, **build_kw)
    build_time = time.time() - t0
    print(f"Build time: {build_time:.2f}s", file=out)
    adapter = VectorIndexerFaissAdapter(idx, k)
    results = {"backend": "vector_indexer", "n": n, "d": d, "nlist": nlist, "k": k, "build_time_s": build_time,
               "search_results": {}}
    for nprobe in nprobes:
        adapter.nprobe = nprobe
        print("nprobe=%-4d" % nprobe, end="  ", file=out)
        results["search_results"][f"nprobe={nprobe}"] = eval_setting(adapter, xq, gt, k, min_time, out)
    return results


def save_results(all_results, output_dir):
    """faiss_bench_results.json + faiss_bench_results.md, laid out like the reference's (bench_all_ivf.py:514-533)."""
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "faiss_bench_results.json"), "w") as f:
        json.dump(all_results, f, indent=2)
    with open(os.path.join(output_dir, "faiss_bench_results.md"), "w") as f:
        f.write("# IVF Benchmark Results (Official Faiss Methodology)\n\n")
        f.write("eval_setting(): every setting runs for at least min_test_duration and reports the mean;\n"
                "R@r = fraction of queries whose true nearest neighbour is among the first r results.\n\n")
        for result in all_results:
            f.write(f"## {result['backend']}\n\n")
            f.write(f"- n={result['n']}, d={result['d']}, nlist={result['nlist']}, k={result['k']}\n")
            f.write(f"- Build time: {result['build_time_s']:.2f}s\n\n")
            f.write("| nprobe | R@1 | R@10 | R@100 | ms/query | QPS |\n")
            f.write("|--------|-----|------|-------|----------|-----|\n")
            for key, res in result["search_results"].items():
                rec = {int(r): v for r, v in res.get("recalls", {}).items()}
                cells = [("%.4f" % rec[r]) if r in rec else "-" for r in RANKS]
                f.write(f"| {key.replace('nprobe=', '')} | {cells[0]} | {cells[1]} | {cells[2]} | "
                        f"{res['ms_per_query']:.3f} | {res['qps']:.1f} |\n")
            f.write("\n")


def main(argv=None):
    ap = argparse.ArgumentParser(description="bench_all_ivf.py-shaped sweep of the B200 vector_indexer")
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--nq", type=int, default=1000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--nprobes", default="1,2,4,8,16,32,64")
    ap.add_argument("--nlist", type=int, default=0, help="0 = the reference heuristic (suggest_nlist)")
    ap.add_argument("--min_test_duration", type=float, default=3.0)
    ap.add_argument("--output-dir", default="./faiss_bench_official_results")
    ap.add_argument("--work-dir", default=None)
    a = ap.parse_args(argv)
    xb, xq = synthetic_dataset(a.n, a.d, a.nq, a.seed)
    gt = exact_ground_truth(xb, xq, a.k)
    res = benchmark_vector_indexer(xb, xq, gt, a.k, [int(p) for p in a.nprobes.split(",")], a.min_test_duration, a.work_dir,
                                   nlist=a.nlist)
    save_results([res], a.output_dir)
    return res


if __name__ == "__main__":
    main()
