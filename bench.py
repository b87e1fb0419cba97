#!/usr/bin/env python
"""bench.py -- headline benchmark of the IVF search hot path on B200.

Metric (BASELINE.json): search QPS at recall@10 >= 0.9 on synthetic SIFT-1M-shaped data
(configs[1]: 1M x 128 fp32, nlist=1024, nq=10k, k=10), plus the list-scan roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference [...]                        the reference's CPU algorithm (oracle port)

One "step" = one search of the whole nq-query batch at the smallest n_probe of the sweep
{1,2,4,...} whose recall@10 (set intersection against exact brute force) is >= 0.9.
`value`  : QPS with queries and index resident in HBM (vidx_search_device), CUDA events.
`e2e`    : QPS through vidx_search with pinned HOST buffers (H2D + D2H inside the timed region).
N > 1    : every rank holds the shards it owns (vidx_set_partition), queries are replicated,
           per-rank top-k are exchanged with one NCCL all-gather and merged on the device;
           strong scaling (total work fixed), max-over-ranks time.
Inputs are larger than L2 (512 MB index vs 126 MB), so no explicit L2 flush between steps.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "vector-indexer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "search QPS at recall@10>=0.9 (SIFT-1M shape, nq=10k)"
UNIT = "queries/s"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4: SMs x lanes x 2 x max SM clock (no FP32 figure in MEASURED_PEAKS)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def workload(args):
    return dict(n=args.n, d=args.d, nq=args.nq, k=args.k, nlist=args.nlist, seed=42)


def gen_data(w):
    """bench/faiss_bench_official/bench_all_ivf.py:67-69"""
    rng = np.random.default_rng(w["seed"])
    xb = rng.standard_normal((w["n"], w["d"])).astype(np.float32)
    xq = rng.standard_normal((w["nq"], w["d"])).astype(np.float32)
    return xb, xq


def recall_at_k(I, gt):
    """tests/test_utils/mod.rs:214-221 (set intersection), averaged over queries."""
    hit = 0
    for a, b in zip(I, gt):
        hit += len(set(a.tolist()) & set(b.tolist()))
    return hit / gt.size


def ncu_traffic(w, nprobe):
    """DRAM bytes of the scan kernel per step from the committed ncu capture of this workload (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            e = json.load(f)
        if all(e["workload"].get(k) == w[k] for k in w) and e["workload"].get("nprobe") == nprobe:
            return e["dram_bytes"]
    return None


def cached_nprobe(w):
    path = os.path.join(ROOT, "profiles", "recall_curve.json")
    if os.path.exists(path):
        with open(path) as f:
            for e in json.load(f):
                if all(e["workload"].get(k) == w[k] for k in w):
                    return e
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------
def cpu_baseline_sample(O, oix, xq, k, nprobe, seconds=12.0):
    """The oracle (port of src/ivf_index.rs:190-267) on every host core, on a bounded
    prefix of the same query batch."""
    threads = O.num_threads()
    done, t0 = 0, time.perf_counter()
    chunk = max(8, 2 * threads)
    while done < len(xq):
        oix.search_batch(xq[done:done + chunk], k, nprobe, nthreads=0)
        done += min(chunk, len(xq) - done)
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {done} of {len(xq)} queries, n_probe={nprobe}, k={k}, oracle with OpenMP over queries "
                      f"(the reference itself runs queries one at a time on one thread)"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm.  The Rust crate cannot be compiled in
    this image (no cargo/rustc), so this is the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    w = workload(args)
    xb, xq = gen_data(w)
    t0 = time.perf_counter()
    oix = O.Ivf.fit(xb, seed=42, nlist=w["nlist"])
    build_s = time.perf_counter() - t0
    cached = cached_nprobe(w)
    if args.nprobe:
        nprobe = args.nprobe
    elif cached:
        nprobe = cached["nprobe"]
    else:  # derive it on a query sample (exact brute force on the CPU is slow)
        s = xq[:200]
        gt = O.brute_force_topk(xb, s, w["k"])
        nprobe = 1
        while nprobe < oix.nlist:
            _, I = oix.search_batch(s, w["k"], nprobe, nthreads=0)
            if recall_at_k(I, gt) >= 0.9:
                break
            nprobe *= 2
    threads = O.num_threads()
    sample = max(threads * 4, 64)
    per_step = []
    for s in range(args.warmup + args.steps):
        q = xq[(s * sample) % (len(xq) - sample):][:sample]
        t0 = time.perf_counter()
        oix.search_batch(q, w["k"], nprobe, nthreads=0)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            per_step.append(dt)
    total = float(np.sum(per_step))
    qps = sample * len(per_step) / total
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(per_step), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 1Mx128 fp32 nlist=1024 nq=10k k=10", **w, "nprobe": nprobe,
                       "step": f"{sample}-query sample of the batch per step", "index_build_s": build_s},
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} queries per step x {len(per_step)} steps, n_probe={nprobe}"},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from vector_indexer_py import _ffi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = workload(args)
    k = w["k"]
    xb, xq = gen_data(w)
    nq, d = xq.shape

    t0 = time.perf_counter()
    ix = _ffi.Index(d, local).build(xb, seed=42, nlist=w["nlist"])
    if args.scan_mode:
        ix.set_scan_mode(args.scan_mode)
    if args.coarse_mode:
        ix.set_coarse_mode(args.coarse_mode)
    build_s = time.perf_counter() - t0
    # The library launches on the stream it is given; torch's legacy default stream has handle 0,
    # which the ABI reads as "use the handle's own stream", so run everything on an explicit
    # torch stream: torch.cuda.Event then brackets exactly the kernels being timed.
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    d_xq = torch.from_numpy(xq).cuda()
    d_D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    d_I = torch.empty((nq, k), dtype=torch.int64, device="cuda")

    def search_dev(nprobe):
        ix.search_device(d_xq.data_ptr(), nq, k, nprobe, d_D.data_ptr(), d_I.data_ptr(), stream)

    # ---- ground truth + n_probe sweep (untimed; single-GPU view of the whole index) ----------
    search_dev(ix.nlist)  # probing every list = exact brute force with the same arithmetic
    torch.cuda.synchronize()
    gt = d_I.cpu().numpy().copy()
    curve, nprobe = [], None
    p = 1
    while True:
        p = min(p, ix.nlist)
        search_dev(p)
        torch.cuda.synchronize()
        r = recall_at_k(d_I.cpu().numpy(), gt)
        curve.append({"nprobe": p, "recall_at_10": r})
        if nprobe is None and r >= 0.9:
            nprobe = p
            if not args.full_curve:
                break
        if p >= ix.nlist:
            break
        p *= 2
    if args.nprobe:
        nprobe = args.nprobe
    recall = next((c["recall_at_10"] for c in curve if c["nprobe"] == nprobe), None)

    # ---- distributed layout -----------------------------------------------------------------
    # index (default, north_star 4): every rank owns part of the index, queries are replicated, the per-rank top-k are
    #   all-gathered and merged.  queries: every rank keeps the whole index (it fits: 0.8 GB) and answers its slice of the
    #   batch; the slices are all-gathered.  Both are strong scaling: the 10 k-query batch is the fixed total work.
    split_queries = world > 1 and args.multi == "queries"
    q_lo, q_hi = 0, nq
    if world > 1 and not split_queries:
        ix.set_partition(rank, world)
        g_D = torch.empty((world, nq, k), dtype=torch.float32, device="cuda")
        g_I = torch.empty((world, nq, k), dtype=torch.int64, device="cuda")
        m_D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        m_I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    if split_queries:
        per = (nq + world - 1) // world
        q_lo, q_hi = min(nq, rank * per), min(nq, (rank + 1) * per)
        s_D = torch.full((per, k), float("inf"), dtype=torch.float32, device="cuda")
        s_I = torch.full((per, k), -1, dtype=torch.int64, device="cuda")
        a_D = torch.empty((world * per, k), dtype=torch.float32, device="cuda")
        a_I = torch.empty((world * per, k), dtype=torch.int64, device="cuda")
        m_D, m_I = a_D[:nq], a_I[:nq]

    def step_device():
        if split_queries:
            if q_hi > q_lo:
                ix.search_device(d_xq[q_lo:q_hi].data_ptr(), q_hi - q_lo, k, nprobe, s_D.data_ptr(), s_I.data_ptr(), stream)
            dist.all_gather_into_tensor(a_D, s_D)
            dist.all_gather_into_tensor(a_I, s_I)
            return
        search_dev(nprobe)
        if world > 1:
            dist.all_gather_into_tensor(g_D, d_D)
            dist.all_gather_into_tensor(g_I, d_I)
            _ffi.merge_topk_device(local, g_D.data_ptr(), g_I.data_ptr(), world, nq, k, m_D.data_ptr(), m_I.data_ptr(), stream)

    h_xq = torch.from_numpy(xq).pin_memory()
    h_D = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    h_I = torch.empty((nq, k), dtype=torch.int64).pin_memory()

    def step_e2e():
        if world == 1:
            ix.search_host_ptr(h_xq.data_ptr(), nq, k, nprobe, h_D.data_ptr(), h_I.data_ptr())
        else:
            d_xq.copy_(h_xq, non_blocking=True)
            step_device()
            h_D.copy_(m_D, non_blocking=True)
            h_I.copy_(m_I, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _ffi.kernel_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _ffi.kernel_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    if args.profile_window:
        # ncu --profile-from-start off: exactly one warmed-up step between cudaProfilerStart/Stop
        for _ in range(args.warmup):
            step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.result()
    # end-to-end: wall clock around host-facing calls (copies included), max over ranks
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # ---- correctness of what was timed ------------------------------------------------------
    final_I = (m_I if world > 1 else d_I).cpu().numpy()
    final_recall = recall_at_k(final_I, gt)

    # ---- per-stage times + roofline of the scan kernel (instrumented pass, not the timed one) ----
    peaks, peak_kind = load_peaks()
    ix.set_profiling(True)
    reps = 5

    def staged(nq_used, nprobe_used):
        acc = None
        for _ in range(reps):
            ix.search_device(d_xq.data_ptr(), nq_used, k, nprobe_used, d_D.data_ptr(), d_I.data_ptr(), stream)
            torch.cuda.synchronize()
            s = ix.stats()
            acc = s if acc is None else {kk: (acc[kk] + s[kk] if kk.startswith("ms_") else s[kk]) for kk in s}
        return {kk: (acc[kk] / reps if kk.startswith("ms_") else acc[kk]) for kk in acc}

    stage = staged(q_hi - q_lo if split_queries else nq, nprobe)
    tc_used = stage["n_tc_items"] > 0
    tc_peak = peaks["bf16_tflops"]  # the filter runs tcgen05 kind::f16 (same rate as bf16): the measured dense peak
    if tc_used:
        t = stage["ms_scan_tc"] / 1e3
        ach = stage["tc_mma_flops"] / t / 1e12
        roofline = {"kernel": "scan_tc_kernel (tcgen05 FP16 pre-filter of the list scan, TMEM accumulators, filter epilogue; "
                              "bounds launch + main launch)",
                    "bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak,
                    "peak_source": f"{peak_kind} bf16 dense GEMM (cuBLAS), burst",
                    "flops_per_launch": stage["tc_mma_flops"], "ms_per_launch": stage["ms_scan_tc"],
                    "pairs_per_launch": stage["tc_mma_flops"] // (2 * d), "filter_survivors": stage["n_tc_survivors"],
                    "queries_redone_exactly": stage["n_tc_overflow"],
                    "reference_arithmetic_equivalent_tflops": stage["scan_flops"] / t / 1e12,
                    "hbm_gbs_algorithmic": stage["scan_bytes_algorithmic"] / t / 1e9,
                    "traffic": ncu_traffic(w, nprobe) if world == 1 else None,
                    "traffic_source": "profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of both launches "
                                      "(ncu --set full); the fp16 shadow store is 256 MB, read once",
                    "note": f"2*D flop per (query, vector) pair, on this rank; at n_probe={nprobe} each probed list is shared by "
                            f"~{stage['n_pairs'] // max(1, ix.nlist)} queries on average, so the contraction, not HBM, bounds the scan; "
                            "the tensor pipe is busy 66 % of the main launch (ncu); one thread issues an N=128 tcgen05.mma only every ~125 cycles (two issuers interleave) and the epilogue (min tree + survivor queue) paces the rest"}
    else:
        t = stage["ms_scan"] / 1e3
        ach = stage["scan_flops"] / t / 1e12
        roofline = {"kernel": "scan_dense_kernel+scan_sparse_kernel (exact FP32 list scan)", "bound": "fp32", "achieved": ach,
                    "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP32_PEAK_TFLOPS,
                    "peak_source": "computed: 148 SM x 128 lanes x 2 x 1.965 GHz", "flops_per_launch": stage["scan_flops"],
                    "ms_per_launch": stage["ms_scan"], "traffic": None}
    # the HBM-bound operating point of the same kernel: one 128-query tile, so every probed list is
    # streamed from HBM exactly once (the latency-oriented small-batch case)
    nq_small = min(128, nq)
    small = staged(nq_small, nprobe)
    ts = (small["ms_scan_tc"] if small["n_tc_items"] > 0 else small["ms_scan"]) / 1e3
    roofline_hbm = {"kernel": "scan_tc_kernel" if small["n_tc_items"] > 0 else "scan_dense/sparse_kernel", "bound": "hbm",
                    "achieved": small["scan_bytes_algorithmic"] / ts / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": small["scan_bytes_algorithmic"] / ts / 1e9 / peaks["hbm_gbs"], "peak_source": peak_kind,
                    "bytes_per_launch": small["scan_bytes_algorithmic"], "ms_per_launch": ts * 1e3, "traffic": None,
                    "workload": f"first {nq_small} queries of the batch, n_probe={nprobe}: distinct probed lists "
                                f"len*(4D+8) + queries + probe lists + outputs",
                    "qps": nq_small / (small["ms_total"] / 1e3)}
    ix.set_profiling(False)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O
        oix = O.Ivf.from_labels(xb, ix.train_centroids(), ix.train_labels())
        cpu = cpu_baseline_sample(O, oix, xq, k, nprobe)
        # the sample doubles as a live parity check of what was timed
        m = 32
        Do, Io = oix.search_batch(xq[:m], k, nprobe, nthreads=0)
        search_dev(nprobe)
        torch.cuda.synchronize()
        assert np.array_equal(d_D[:m].cpu().numpy().view(np.uint32), Do.view(np.uint32)), "GPU distances differ from the oracle"
        assert np.array_equal(d_I[:m].cpu().numpy(), Io), "GPU ids differ from the oracle"

    if rank == 0:
        qps = nq * args.steps / (ms_dev / 1e3)
        line = {"metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: 1Mx128 fp32 nlist=1024 nq=10k k=10", **w, "nprobe": nprobe,
                           "recall_at_10": final_recall, "recall_curve": curve, "nlist_nonempty": ix.nlist,
                           "l2": "inputs larger than L2 (index 512 MB)", "index_build_s": build_s,
                           "parallelism": ("1 GPU" if world == 1 else
                                           f"index replicated, batch split over {world} GPUs, NCCL all-gather of the slices" if split_queries
                                           else f"{ix.partition_kind} over {world} GPUs, queries replicated, NCCL all-gather of per-GPU "
                                                f"top-k + device merge")},
                "e2e": {"value": nq * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(xq.nbytes),
                        "d2h_bytes_per_step": int(nq * k * 12), "ms_per_step": 1e3 * e2e_s / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm,
                "stages_ms": {kk: stage[kk] for kk in stage if kk.startswith("ms_")},
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nlist", type=int, default=1024)
    ap.add_argument("--nprobe", type=int, default=0, help="0 = smallest power of two with recall@10 >= 0.9")
    ap.add_argument("--full-curve", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi", default="index", choices=["index", "queries"],
                    help="N > 1: partition the index over the GPUs (default, north_star) or replicate it and split the batch")
    ap.add_argument("--scan-mode", type=int, default=0, help="experiments: vidx_set_scan_mode (0 = auto, what the bench line is quoted on)")
    ap.add_argument("--coarse-mode", type=int, default=0, help="experiments: vidx_set_coarse_mode (0 = auto)")
    ap.add_argument("--profile-window", action="store_true",
                    help="bracket one warmed-up step with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
