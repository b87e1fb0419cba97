#!/usr/bin/env python
"""bench.py -- headline benchmark of the IVF search hot path on B200.

Metric (BASELINE.json): search QPS at recall@10 >= 0.9 on synthetic SIFT-1M-shaped data
(configs[1]: 1M x 128 fp32, nlist=1024, nq=10k, k=10), plus the list-scan roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference [...]                        the reference's CPU algorithm (oracle port)
  python bench.py --config c4 --gpus N                          another BASELINE config as the timed workload

One "step" = one search of the whole nq-query batch at the smallest n_probe of the sweep
{1,2,4,...} whose recall@10 (set intersection against an independent float64 brute force) is >= 0.9.
`value`  : QPS with queries and index resident in HBM (vidx_search_device / vidx_search_multi_device), CUDA events.
`e2e`    : QPS through vidx_search / vidx_search_multi with pinned HOST buffers (H2D + D2H inside the timed region).
N > 1    : one process per GPU.  Every rank declares its partition BEFORE the build (vidx_set_partition), so only the part
           it owns ever reaches its HBM; queries are replicated; the exchange (probe lists, then ONE packed all-gather of the
           per-rank top-k + device merge) runs inside the library over its own NCCL communicator (vidx_comm_init).  Strong
           scaling (the 10 k-query batch is the fixed total work), max-over-ranks time.  The line also carries `c4`: the
           10M x 128 sharded config (BASELINE configs[3]) measured the same way on the same ranks.
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

# BASELINE.json configs: (n, d, nlist (0 = the reference heuristic, utils.rs:9-16), default n_probe (0 = from the recall sweep))
CONFIGS = {
    "c1": dict(n=50_000, d=64, nlist=0, nprobe=0, name="configs[0]: 50kx64 fp32 nlist=448 (heuristic) nq=10k k=10"),
    "c2": dict(n=1_000_000, d=128, nlist=1024, nprobe=0, name="configs[1]: 1Mx128 fp32 nlist=1024 nq=10k k=10"),
    "c4": dict(n=10_000_000, d=128, nlist=0, nprobe=32, name="configs[3]: 10Mx128 fp32 nlist=12652 (heuristic) n_probe=32 nq=10k k=10"),
    "c5": dict(n=100_000_000, d=96, nlist=65536, nprobe=32, name="configs[4]: 100Mx96 fp32 nlist=65536 n_probe=32 nq=10k k=10"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def workload(args):
    return dict(n=args.n, d=args.d, nq=args.nq, k=args.k, nlist=args.nlist, seed=42)


def gen_data(w, chunk=1_000_000):
    """bench/faiss_bench_official/bench_all_ivf.py:67-69: xb then xq from ONE default_rng(seed) stream (drawn in chunks:
    the stream is the same as one big call)."""
    rng = np.random.default_rng(w["seed"])
    xb = np.empty((w["n"], w["d"]), np.float32)
    for i0 in range(0, w["n"], chunk):
        i1 = min(w["n"], i0 + chunk)
        xb[i0:i1] = rng.standard_normal((i1 - i0, w["d"]))
    xq = rng.standard_normal((w["nq"], w["d"])).astype(np.float32)
    return xb, xq


def recall_at_k(I, gt):
    """tests/test_utils/mod.rs:214-221 (set intersection), averaged over queries."""
    hit = 0
    for a, b in zip(I, gt):
        hit += len(set(a.tolist()) & set(b.tolist()))
    return hit / gt.size


def faiss_recalls(I, gt):
    """bench_all_ivf.py:283-363 (the Faiss convention): R@r = share of queries whose TRUE nearest neighbour is among the
    first r results."""
    nn = gt[:, :1]
    return {f"R@{r}": float((I[:, :r] == nn).any(axis=1).mean()) for r in (1, 10) if r <= I.shape[1]}


def curve_point(p, I, gt):
    return {"nprobe": p, "recall_at_10": recall_at_k(I, gt), **faiss_recalls(I, gt)}


def brute_force_topk_f64(xb, xq, k, device="cuda", block=32768):
    """Independent ground truth (replaces faiss IndexFlatL2 of bench_all_ivf.py:75-78): exact float64 squared L2 by
    blocks with torch on the GPU -- none of the library's code.  xb may be a host array or a device tensor.
    Returns (D float64 [nq, k], I int64 [nq, k])."""
    import torch
    q = torch.as_tensor(xq, device=device).double()
    qn = (q * q).sum(1, keepdim=True)
    best_d = torch.full((len(q), k), float("inf"), dtype=torch.float64, device=device)
    best_i = torch.full((len(q), k), -1, dtype=torch.int64, device=device)
    n = len(xb)
    for b0 in range(0, n, block):
        b1 = min(n, b0 + block)
        x = torch.as_tensor(xb[b0:b1], device=device).double()
        d = qn - 2.0 * (q @ x.T) + (x * x).sum(1)[None, :]
        kk = min(k, b1 - b0)
        dv, di = torch.topk(d, kk, dim=1, largest=False)
        cat_d = torch.cat([best_d, dv], 1)
        cat_i = torch.cat([best_i, di + b0], 1)
        o = torch.argsort(cat_d, dim=1, stable=True)[:, :k]
        best_d, best_i = torch.gather(cat_d, 1, o), torch.gather(cat_i, 1, o)
    return best_d.cpu().numpy(), best_i.cpu().numpy()


def ncu_traffic(w, nprobe, key="dram_bytes"):
    """DRAM bytes of the scan kernel per step from the committed ncu capture of this workload (profiles/ncu_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            e = json.load(f)
        if all(e["workload"].get(k) == w[k] for k in w) and e["workload"].get("nprobe") == nprobe:
            return e.get(key)
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
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle port of src/ivf_index.rs:190-267
# ----------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_threads_setup():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm states its own thread count instead of inheriting that."""
    n = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    import oracle as O
    O.lib()
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except Exception:
        pass
    return O, n


CPU_CHUNK = 64  # queries per oracle call in every CPU measurement (the same in cpu_baseline and --impl reference)


def cpu_search_qps(oix, xq, k, nprobe, nthreads, seconds, start=0):
    """Chunks of CPU_CHUNK queries from `start` until `seconds` have passed.  nthreads = 1 is the reference's own
    behaviour (one query at a time on one thread, bindings/python/src/lib.rs:74-97); nthreads = all cores is the generous
    variant (OpenMP over the queries of a chunk)."""
    done, t0 = 0, time.perf_counter()
    while True:
        a = (start + done) % max(1, len(xq) - CPU_CHUNK)
        oix.search_batch(xq[a:a + CPU_CHUNK], k, nprobe, nthreads=nthreads)
        done += CPU_CHUNK
        if time.perf_counter() - t0 > seconds:
            break
    return done / (time.perf_counter() - t0), done


def cpu_baseline_block(O, ncores, oix, xq, k, nprobe, seconds_all=10.0, seconds_one=6.0):
    allc, n_all = cpu_search_qps(oix, xq, k, nprobe, ncores, seconds_all)
    one, n_one = cpu_search_qps(oix, xq, k, nprobe, 1, seconds_one)
    return {"value": allc, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": f"first {n_all} queries of the batch in chunks of {CPU_CHUNK}, n_probe={nprobe}, k={k}: oracle port with "
                      f"OpenMP over the queries of a chunk on {ncores} threads",
            "faithful": {"value": one, "unit": UNIT, "cores": 1,
                         "sample": f"first {n_one} queries, one query at a time on one thread -- what the reference's Python "
                                   f"entry point does (bindings/python/src/lib.rs:74-97), minus its per-query shard file reads"}}


def config_dict(args, w, nprobe):
    """The workload, identical in both arms (the driver compares it key by key)."""
    return {"workload": CONFIGS[args.config]["name"] if args.config in CONFIGS else "custom", **w, "nprobe": nprobe}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm.  The Rust crate cannot be compiled in
    this image (no cargo/rustc), so this is the oracle port on all host threads (stated in `cores`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, ncores = cpu_threads_setup()
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
            _, I = oix.search_batch(s, w["k"], nprobe, nthreads=ncores)
            if recall_at_k(I, gt) >= 0.9:
                break
            nprobe *= 2
    per_step = []
    for s in range(args.warmup + args.steps):
        a = (s * CPU_CHUNK) % (len(xq) - CPU_CHUNK)
        t0 = time.perf_counter()
        oix.search_batch(xq[a:a + CPU_CHUNK], w["k"], nprobe, nthreads=ncores)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            per_step.append(dt)
    total = float(np.sum(per_step))
    qps = CPU_CHUNK * len(per_step) / total
    one, n_one = cpu_search_qps(oix, xq, w["k"], nprobe, 1, 6.0)
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(per_step), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, w, nprobe),
            "details": {"step": f"one chunk of {CPU_CHUNK} queries of the batch per step (bounded sample of the workload)",
                        "index_build_s": build_s, "nlist_nonempty": oix.nlist},
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": ncores, "kind": "port",
                             "sample": f"{CPU_CHUNK} queries per step x {len(per_step)} steps, n_probe={nprobe}, OpenMP over the "
                                       f"queries of a chunk on {ncores} threads",
                             "faithful": {"value": one, "unit": UNIT, "cores": 1,
                                          "sample": f"first {n_one} queries, one at a time on one thread "
                                                    f"(bindings/python/src/lib.rs:74-97)"}},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
class Harness:
    """One process per GPU: device, stream, (optionally) torch.distributed for the host-side plumbing only -- the
    data-path collectives are the library's own (vidx_comm_init)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        # The library launches on the stream it is given; torch's legacy default stream has handle 0,
        # which the ABI reads as "use a stream of the handle", so run everything on an explicit
        # torch stream: torch.cuda.Event then brackets exactly the kernels being timed.
        self.tstream = torch.cuda.Stream()
        torch.cuda.set_stream(self.tstream)
        self.stream = self.tstream.cuda_stream
        assert self.stream != 0

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def comm_for(self, ix):
        """Rank 0's NCCL id travels over the host-side process group; the communicator itself belongs to the library."""
        from vector_indexer_py import _ffi
        obj = [_ffi.comm_unique_id() if self.rank == 0 else None]
        self.dist.broadcast_object_list(obj, src=0)
        ix.comm_init(self.rank, self.world, obj[0])

    def timed(self, fn, steps, warmup):
        from vector_indexer_py import _ffi
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _ffi.kernel_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        return ms, _ffi.kernel_launch_count() - l0

    def timed_wall(self, fn, steps, warmup):
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.barrier()
        return self.max_over_ranks(time.perf_counter() - t0)


class Searcher:
    """A built index + device / pinned-host buffers for one batch; search through the same entry points at any N."""

    def __init__(self, H, w, xb, xq, nlist, scan_mode=0, coarse_mode=0, d_xb=None, parts=0):
        from vector_indexer_py import _ffi
        torch = H.torch
        self.H, self.k, self.nq, self.d = H, w["k"], len(xq), xq.shape[1]
        t0 = time.perf_counter()
        ix = _ffi.Index(self.d, H.local)
        # The ranks form a grid of `parts` index parts x world / parts query groups (vidx_search_multi): parts = world is the
        # plain sharded index, parts = 1 a replica per GPU with the batch split by query.
        self.parts = parts if parts and H.world % parts == 0 else H.world
        if self.parts > 1:
            ix.set_partition(H.rank % self.parts, self.parts)   # BEFORE the build: only the owned part reaches HBM
        if d_xb is not None:
            ix.build_device(d_xb.data_ptr(), len(d_xb), seed=42, nlist=nlist)
        else:
            ix.build(xb, seed=42, nlist=nlist)
        self.build_s = time.perf_counter() - t0
        if H.world > 1:
            H.comm_for(ix)
        if scan_mode:
            ix.set_scan_mode(scan_mode)
        if coarse_mode:
            ix.set_coarse_mode(coarse_mode)
        self.ix = ix
        self.d_xq = torch.from_numpy(xq).cuda()
        self.d_D = torch.empty((self.nq, self.k), dtype=torch.float32, device="cuda")
        self.d_I = torch.empty((self.nq, self.k), dtype=torch.int64, device="cuda")
        self.h_xq = torch.from_numpy(xq).pin_memory()
        self.h_D = torch.empty((self.nq, self.k), dtype=torch.float32).pin_memory()
        self.h_I = torch.empty((self.nq, self.k), dtype=torch.int64).pin_memory()

    def search_dev(self, nprobe, nq=None):
        nq = self.nq if nq is None else nq
        a = (self.d_xq.data_ptr(), nq, self.k, nprobe, self.d_D.data_ptr(), self.d_I.data_ptr(), self.H.stream)
        if self.H.world > 1:
            self.ix.search_multi_device(*a)
        else:
            self.ix.search_device(*a)

    def search_e2e(self, nprobe):
        a = (self.h_xq.data_ptr(), self.nq, self.k, nprobe, self.h_D.data_ptr(), self.h_I.data_ptr())
        if self.H.world > 1:
            self.ix.search_multi_host_ptr(*a)
        else:
            self.ix.search_host_ptr(*a)

    def result_ids(self):
        self.H.torch.cuda.synchronize()
        return self.d_I.cpu().numpy()

    def staged(self, nq_used, nprobe, reps=5):
        """Per-stage CUDA-event times (instrumented passes, not the timed ones), averaged."""
        self.ix.set_profiling(True)
        self.search_dev(nprobe, nq_used)  # untimed: buffers of a stage that has not run yet on this handle get allocated here
        self.H.torch.cuda.synchronize()
        acc = None
        for _ in range(reps):
            self.search_dev(nprobe, nq_used)
            self.H.torch.cuda.synchronize()
            s = self.ix.stats()
            acc = s if acc is None else {kk: (acc[kk] + s[kk] if kk.startswith("ms_") else s[kk]) for kk in s}
        self.ix.set_profiling(False)
        return {kk: (acc[kk] / reps if kk.startswith("ms_") else acc[kk]) for kk in acc}


def recall_sweep(S, gt, gt_rows, full_curve, forced):
    """Smallest n_probe of {1, 2, 4, ...} with recall@10 >= 0.9 against the independent ground truth (rows gt_rows of the batch)."""
    curve, nprobe, p = [], None, 1
    if forced and not full_curve:  # a config that fixes n_probe (BASELINE configs[3], [4]): measure the recall there only
        S.search_dev(min(forced, S.ix.nlist))
        return forced, [curve_point(forced, S.result_ids()[gt_rows], gt)]
    while True:
        p = min(p, S.ix.nlist)
        S.search_dev(p)
        curve.append(curve_point(p, S.result_ids()[gt_rows], gt))
        r = curve[-1]["recall_at_10"]
        if nprobe is None and r >= 0.9:
            nprobe = p
            if not full_curve:
                break
        if p >= S.ix.nlist:
            break
        p *= 2
    if forced:
        nprobe = forced
    return nprobe or curve[-1]["nprobe"], curve


def measure(S, nprobe, steps, warmup):
    H = S.H
    sampler = ClockSampler(H.local)
    sampler.start()
    ms_dev, launches = H.timed(lambda: S.search_dev(nprobe), steps, warmup)
    clocks = sampler.result()
    e2e_s = H.timed_wall(lambda: S.search_e2e(nprobe), steps, warmup)
    return ms_dev, launches, clocks, e2e_s


def scan_rooflines(S, w, nprobe, peaks, peak_kind):
    H, nq, d = S.H, S.nq, S.d
    stage = S.staged(nq, nprobe)
    tc_used = stage["n_tc_items"] > 0
    tc_peak = peaks["bf16_tflops"]  # the filter runs tcgen05 kind::f16 (same rate as bf16): the measured dense peak
    if tc_used:
        t = stage["ms_scan_tc"] / 1e3
        ach = stage["tc_mma_flops"] / t / 1e12
        roofline = {"kernel": "scan_tc_kernel (tcgen05 FP16 pre-filter of the list scan, TMEM accumulators, filter epilogue; "
                              "bounds launch + main launch)",
                    "bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak,
                    "peak_source": f"{peak_kind} bf16 dense GEMM (cuBLAS), burst",
                    "peak_sustained": peaks.get("bf16_tflops_sustained"),
                    "frac_of_sustained": (ach / peaks["bf16_tflops_sustained"]) if peaks.get("bf16_tflops_sustained") else None,
                    "flops_per_launch": stage["tc_mma_flops"], "ms_per_launch": stage["ms_scan_tc"],
                    "pairs_per_launch": stage["tc_mma_flops"] // (2 * d), "filter_survivors": stage["n_tc_survivors"],
                    "queries_redone_exactly": stage["n_tc_overflow"],
                    "reference_arithmetic_equivalent_tflops": stage["scan_flops"] / t / 1e12,
                    "hbm_gbs_algorithmic": stage["scan_bytes_algorithmic"] / t / 1e9,
                    "traffic": ncu_traffic(w, nprobe) if H.world == 1 else None,
                    "traffic_source": "profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of both launches "
                                      "(ncu --set full); the fp16 shadow store is 256 MB, read once",
                    "note": f"2*D flop per (query, vector) pair, on this rank; at n_probe={nprobe} each probed list is shared by "
                            f"~{stage['n_pairs'] // max(1, S.ix.nlist)} queries on average, so the contraction, not HBM, bounds the scan"}
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
    small = S.staged(nq_small, nprobe)
    ts = (small["ms_scan_tc"] if small["n_tc_items"] > 0 else small["ms_scan"]) / 1e3
    roofline_hbm = {"kernel": "scan_tc_kernel" if small["n_tc_items"] > 0 else "scan_dense/sparse_kernel", "bound": "hbm",
                    "achieved": small["scan_bytes_algorithmic"] / ts / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": small["scan_bytes_algorithmic"] / ts / 1e9 / peaks["hbm_gbs"], "peak_source": peak_kind,
                    "bytes_per_launch": small["scan_bytes_algorithmic"], "ms_per_launch": ts * 1e3,
                    # what the filter actually streams: the fp16 shadow store (16 B per 8 dims + 16 B of norm terms per vector)
                    # instead of the reference's fp32 record -- which is why `achieved` can exceed the copy bandwidth on indexes
                    # whose probed lists are large
                    "bytes_streamed_model": int(small["scan_bytes_algorithmic"] / (4 * d + 8) * (32 * ((d + 15) // 16) + 16)),
                    "traffic": ncu_traffic(w, nprobe, "dram_bytes_nq128") if H.world == 1 else None,
                    "workload": f"first {nq_small} queries of the batch, n_probe={nprobe}: distinct probed lists "
                                f"len*(4D+8) + queries + probe lists + outputs",
                    "qps": nq_small / (small["ms_total"] / 1e3)}
    return stage, roofline, roofline_hbm


def coarse_roofline(S, nprobe, peaks, peak_kind):
    """Coarse quantization both ways: exact FP32 kernel (default) and tcgen05 filter + exact re-check."""
    out = {}
    nq, d = S.nq, S.d
    for mode, name in ((1, "fp32_exact"), (2, "tensor_core_filter")):
        S.ix.set_coarse_mode(mode)
        st = S.staged(nq, nprobe, reps=3)
        ms = st["ms_coarse"] + st["ms_select"]
        alg = 2.0 * d * nq * S.ix.nlist  # 2 * nq * nlist * D
        e = {"ms": ms, "algorithmic_gflop": alg / 1e9, "tflops_algorithmic": alg / (ms / 1e3) / 1e12}
        if mode == 1:
            e.update(frac_fp32_peak_3D=(1.5 * alg) / (st["ms_coarse"] / 1e3) / 1e12 / FP32_PEAK_TFLOPS,
                     kernel="coarse_dist_kernel + select", ms_distances=st["ms_coarse"], ms_select=st["ms_select"])
        else:
            # one extra K-step per tile carries the norms: 2 * 8 * Dh halfs per pair are issued, twice (bounds + main pass)
            dh16 = 16 * ((d + 15) // 16)
            issued = 2.0 * 2.0 * (dh16 + 16) * nq * (128 * ((S.ix.nlist + 127) // 128))
            e.update(frac_tensor_peak_issued=issued / (ms / 1e3) / 1e12 / peaks["bf16_tflops"],
                     kernel="scan_tc_kernel over the centroid table (bounds + frozen pass) + finalize")
        out[name] = e
    S.ix.set_coarse_mode(0)
    st = S.staged(nq, nprobe, reps=3)
    out["auto"] = {"ms": st["ms_coarse"] + st["ms_select"]}
    out["peak_source"] = f"{peak_kind} bf16 burst {peaks['bf16_tflops']} TF/s; FP32 {FP32_PEAK_TFLOPS:.1f} TF/s computed"
    return out


def kmeans_block(xb, ncores, with_oracle):
    """BASELINE configs[2]: mini-batch k-means 1M x 128, k = 4096, 20 iterations, hierarchical final assignment."""
    from vector_indexer_py import _ffi
    runs = []
    for _ in range(3):  # the first run pays the workspace and pinned-buffer allocations: reported, not the headline
        _ffi.kmeans_last_profile()
        t0 = time.perf_counter()
        c, labels, iters = _ffi.kmeans_mini_batch(xb, 4096, 20, seed=42)
        runs.append((time.perf_counter() - t0,) + tuple(_ffi.kmeans_last_profile()))
    gpu_s, host_rng_s, device_wait_s = [sum(r[i] for r in runs[1:]) / 2 for i in range(3)]
    out = {"workload": "configs[2]: mini-batch k-means 1Mx128 k=4096, 20 iterations + hierarchical assignment of all points",
           "gpu_seconds": gpu_s, "gpu_seconds_first_call": runs[0][0], "runs": "3 calls; gpu_seconds = mean of calls 2 and 3",
           "includes": "H2D of the data set (512 MB) and D2H of centroids + labels", "iterations": iters,
           "split_seconds": {"host_serial_random_stream": host_rng_s, "blocked_on_device": device_wait_s,
                             "other_host_and_copies": max(0.0, gpu_s - host_rng_s - device_wait_s)},
           "split_note": "the reference's semantics keep the draws of 20 Fisher-Yates shuffles of 1M indices (kmeans.rs:722-726) and 4095 "
                         "sequential 50 000-term prefix sums (k-means++, :270-287) serial on the host; the device does every distance / "
                         "argmin / mean"}
    if with_oracle:
        import oracle as O
        t0 = time.perf_counter()
        oc, ol, _ = O.kmeans_mini_batch(xb, 4096, 20, seed=42)
        out.update(cpu_seconds=time.perf_counter() - t0, cpu_cores=ncores, cpu_kind="port",
                   centroids_bit_equal=bool(np.array_equal(c.view(np.uint32), oc.view(np.uint32))),
                   labels_equal=bool(np.array_equal(labels, ol)))
    return out


def run_config(H, args, cfg_name, w, nprobe_forced, steps, warmup, rich, parts=0):
    """Build + sweep + timed steps for one workload; `rich` adds rooflines / CPU legs (the headline config)."""
    torch = H.torch
    cfg = CONFIGS.get(cfg_name, {})
    device_gen = cfg_name == "c5"
    if device_gen:
        # too large for a host copy per rank: the same counter-based stream on every GPU (torch Philox, seed 42)
        g = torch.Generator(device="cuda")
        g.manual_seed(42)
        d_xb = torch.empty((w["n"], w["d"]), dtype=torch.float32, device="cuda")
        for i0 in range(0, w["n"], 4_000_000):
            i1 = min(w["n"], i0 + 4_000_000)
            d_xb[i0:i1] = torch.randn((i1 - i0, w["d"]), generator=g, device="cuda")
        xq = torch.randn((w["nq"], w["d"]), generator=g, device="cuda").cpu().numpy()
        xb = None
        torch.cuda.synchronize()
    else:
        xb, xq = gen_data(w)
        d_xb = None
    S = Searcher(H, w, xb, xq, w["nlist"], args.scan_mode, args.coarse_mode, d_xb=d_xb, parts=parts)
    # ---- independent ground truth (float64 brute force, none of the library's code) on a sample of the batch ----
    ngt = w["nq"] if w["n"] <= 2_000_000 else 1000
    gt_rows = np.arange(ngt)
    src = d_xb if device_gen else xb
    _, gt = brute_force_topk_f64(src, xq[:ngt], w["k"])
    del d_xb
    nprobe, curve = recall_sweep(S, gt, gt_rows, args.full_curve, nprobe_forced)
    recall = next((c["recall_at_10"] for c in curve if c["nprobe"] == nprobe), None)
    if args.profile_window and rich:
        # ncu --profile-from-start off: exactly one warmed-up step between cudaProfilerStart/Stop
        for _ in range(warmup):
            S.search_dev(nprobe, args.profile_nq or None)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        S.search_dev(nprobe, args.profile_nq or None)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    ms_dev, launches, clocks, e2e_s = measure(S, nprobe, steps, warmup)
    S.search_dev(nprobe)
    final_recall = recall_at_k(S.result_ids()[gt_rows], gt)
    if recall is None:
        recall = final_recall
    nq, k = w["nq"], w["k"]
    res = {"S": S, "xb": xb, "xq": xq, "nprobe": nprobe, "curve": curve, "recall": final_recall, "ms_dev": ms_dev,
           "launches": launches, "clocks": clocks, "e2e_s": e2e_s,
           "qps": nq * steps / (ms_dev / 1e3), "e2e_qps": nq * steps / e2e_s}
    ix = S.ix
    res["residency"] = {"resident_vectors": int(ix.resident_vectors), "ntotal": int(ix.ntotal),
                        "resident_bytes": int(ix.resident_bytes),
                        "fraction": ix.resident_vectors / max(1, ix.ntotal)}
    groups = H.world // S.parts
    held = f"each rank holds {100.0 * ix.resident_vectors / max(1, ix.ntotal):.1f} % of the vectors in HBM"
    plumbing = (f"coarse stage split by query, {ix.comm_version if H.world > 1 else ''} all-gather of probe lists and of the packed "
                f"per-GPU top-k + device merge, all inside the library (vidx_search_multi)")
    if H.world == 1:
        res["parallelism"] = "1 GPU"
    elif groups == 1:
        res["parallelism"] = f"{ix.partition_kind} over {H.world} GPUs: {held}, queries replicated, {plumbing}"
    elif S.parts == 1:
        res["parallelism"] = (f"{H.world} replicas ({held}), the batch split by query into {groups} groups, {plumbing}")
    else:
        res["parallelism"] = (f"grid of {S.parts} index parts ({ix.partition_kind}) x {groups} query groups over {H.world} GPUs: {held}, "
                              f"{plumbing}")
    res["grid"] = {"index_parts": S.parts, "query_groups": groups}
    return res


def run_ours(args):
    H = Harness()
    torch = H.torch
    w = workload(args)
    peaks, peak_kind = load_peaks()
    # N > 1: the 0.5 GB index of configs[1] is replicated and the batch split by query (a 1 x N grid); the larger
    # configurations are sharded N ways.  --parts overrides; the fully sharded configs[1] numbers go under `c2_sharded`.
    parts = args.parts if args.parts > 0 else (1 if args.config == "c2" else H.world)
    if H.world % parts != 0:
        parts = H.world
    R = run_config(H, args, args.config, w, args.nprobe, args.steps, args.warmup, rich=True, parts=parts)
    S, xb, xq, nprobe = R["S"], R["xb"], R["xq"], R["nprobe"]
    nq, k = w["nq"], w["k"]
    stage, roofline, roofline_hbm = scan_rooflines(S, w, nprobe, peaks, peak_kind)
    coarse = coarse_roofline(S, nprobe, peaks, peak_kind) if not args.lean else None

    cpu, kmeans, parity = None, None, None
    if H.rank == 0 and H.world == 1 and not args.no_cpu_baseline and xb is not None:
        O, ncores = cpu_threads_setup()
        oix = O.Ivf.from_labels(xb, S.ix.train_centroids(), S.ix.train_labels())
        cpu = cpu_baseline_block(O, ncores, oix, xq, k, nprobe)
        # the sample doubles as a live parity check of what was timed
        m = 32
        Do, Io = oix.search_batch(xq[:m], k, nprobe, nthreads=ncores)
        S.search_dev(nprobe)
        torch.cuda.synchronize()
        assert np.array_equal(S.d_D[:m].cpu().numpy().view(np.uint32), Do.view(np.uint32)), "GPU distances differ from the oracle"
        assert np.array_equal(S.d_I[:m].cpu().numpy(), Io), "GPU ids differ from the oracle"
        parity = f"first {m} queries at n_probe={nprobe}: distances bit-identical, ids identical to the oracle port"
        del oix
        if args.config == "c2" and not args.lean:
            kmeans = kmeans_block(xb, ncores, with_oracle=True)

    # ---- configs[1] again, sharded N ways like the reference's own multi-shard search (N > 1) ---------------------------
    c2_sharded = None
    if H.world > 1 and parts != H.world and args.extra != "none":
        del S
        R["S"] = None
        S = None
        torch.cuda.empty_cache()
        try:
            n2 = max(5, args.steps // 2)
            R2 = run_config(H, args, args.config, w, nprobe, n2, 3, rich=False, parts=H.world)
            c2_sharded = {"value": R2["qps"], "unit": UNIT, "ms_per_step": R2["ms_dev"] / n2, "e2e": R2["e2e_qps"],
                          "nprobe": R2["nprobe"], "recall_at_10": R2["recall"], "residency": R2["residency"],
                          "parallelism": R2["parallelism"], "grid": R2["grid"], "gpu_launches": int(R2["launches"])}
            R2["S"] = None
            del R2
        except Exception as e:
            c2_sharded = {"error": f"{type(e).__name__}: {e}"}

    # ---- the sharded 10M config on the same ranks (N > 1, or on request) -------------------------------------------
    c4 = None
    want_c4 = args.extra == "c4" or (args.extra == "auto" and H.world > 1 and args.config == "c2")
    if want_c4:
        S = None
        R["S"] = None
        torch.cuda.empty_cache()
        try:
            c = CONFIGS["c4"]
            w4 = dict(n=c["n"], d=c["d"], nq=args.nq, k=args.k, nlist=c["nlist"], seed=42)
            t0 = time.perf_counter()
            R4 = run_config(H, args, "c4", w4, c["nprobe"], max(5, args.steps // 2), 3, rich=False, parts=H.world)
            c4 = {"workload": c["name"], "n_gpus": H.world, "value": R4["qps"], "unit": UNIT,
                  "ms_per_step": R4["ms_dev"] / max(5, args.steps // 2), "e2e": R4["e2e_qps"], "nprobe": R4["nprobe"],
                  "recall_at_10": R4["recall"], "recall_sample": "first 1000 queries vs float64 brute force",
                  "nlist_nonempty": int(R4["S"].ix.nlist), "num_shards": int(R4["S"].ix.num_shards),
                  "index_build_s": R4["S"].build_s, "residency": R4["residency"], "parallelism": R4["parallelism"], "grid": R4["grid"],
                  "gpu_launches": int(R4["launches"]), "seconds_total": time.perf_counter() - t0}
            R4["S"] = None
        except Exception as e:  # the headline line must survive a failure of the extra workload
            c4 = {"error": f"{type(e).__name__}: {e}"}

    if H.rank == 0:
        line = {"metric": METRIC, "value": R["qps"], "unit": UNIT, "n_gpus": H.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": R["ms_dev"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, w, nprobe),
                "details": {"recall_at_10": R["recall"], "recall_curve": R["curve"],
                            "ground_truth": "float64 brute force (torch, blocks of 32768 rows), independent of the library",
                            "l2": "inputs larger than L2 (index 512 MB)",
                            "parallelism": R["parallelism"], "grid": R["grid"], "residency": R["residency"], "parity": parity,
                            "graph_replay": ("off (VIDX_GRAPH=0)" if os.environ.get("VIDX_GRAPH", "1") == "0" else
                                             "on: the timed steps repeat one search, which the library replays as a CUDA graph from the "
                                             "third call (gpu_launches counts the kernels of the replayed graphs)")},
                "e2e": {"value": R["e2e_qps"], "unit": UNIT, "h2d_bytes_per_step": int(xq.nbytes // R["grid"]["query_groups"]),
                        "d2h_bytes_per_step": int(nq * k * 12), "ms_per_step": 1e3 * R["e2e_s"] / args.steps},
                "gpu_launches": int(R["launches"]), "clocks": R["clocks"], "roofline": roofline, "roofline_hbm": roofline_hbm,
                "roofline_coarse": coarse,
                "stages_ms": {kk: stage[kk] for kk in stage if kk.startswith("ms_")},
                "cpu_baseline": cpu, "kmeans": kmeans, "c2_sharded": c2_sharded, "c4": c4}
        print(json.dumps(line))
    if H.world > 1:
        H.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE config that is built and timed (default: the headline)")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--d", type=int, default=0)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nlist", type=int, default=-1)
    ap.add_argument("--nprobe", type=int, default=-1, help="0 = smallest power of two with recall@10 >= 0.9")
    ap.add_argument("--full-curve", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lean", action="store_true", help="skip the coarse-stage comparison and the k-means block")
    ap.add_argument("--extra", default="auto", choices=["auto", "c4", "none"],
                    help="auto: N > 1 also measures the sharded 10M config (BASELINE configs[3]) and reports it under `c4`")
    ap.add_argument("--parts", type=int, default=0,
                    help="N > 1: index parts of the rank grid (N / parts query groups); 0 = 1 for configs[1] (replicas), N otherwise")
    ap.add_argument("--scan-mode", type=int, default=0, help="experiments: vidx_set_scan_mode (0 = auto, what the bench line is quoted on)")
    ap.add_argument("--coarse-mode", type=int, default=0, help="experiments: vidx_set_coarse_mode (0 = auto)")
    ap.add_argument("--profile-window", action="store_true",
                    help="bracket one warmed-up step with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--profile-nq", type=int, default=0, help="queries of the profiled step (0 = the whole batch)")
    args = ap.parse_args()
    c = CONFIGS[args.config]
    args.n = args.n or c["n"]
    args.d = args.d or c["d"]
    args.nlist = c["nlist"] if args.nlist < 0 else args.nlist
    args.nprobe = c["nprobe"] if args.nprobe < 0 else args.nprobe
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
