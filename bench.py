#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: sampled edges/s and gathered-feature GB/s.

One "step" = one mini-batch through the hot path: k-hop neighbour sampling (quiver.pyg.GraphSageSampler.sample) followed
by the feature gather of the sampled nodes (quiver.Feature.__getitem__) -- the two calls of the reference's training
loop (examples/pyg/reddit_quiver.py:116-122) and of its benchmarks (benchmarks/sample/bench_sampler.py:35-46,
benchmarks/feature/bench_feature.py:36-46, whose metric definitions are reused: SEPS counts adj.edge_index columns,
feature bandwidth counts OUTPUT bytes only).

Workload at N=1 = BASELINE.json configs[1]: ogbn-products-shaped synthetic CSR (2,449,029 nodes, pareto(2) degrees with
mean ~50.5 => ~124 M edges), 1024 seeds per step, fan-out [15,10,5], 100-d fp32 features, feature table fully in HBM.
At N>1 the same graph is replicated per rank (the sampler does not shard: SURVEY.md 8(e)) and the feature table is
row-sharded over the N GPUs and read one-sidedly over NVLink; every rank runs its own batches (weak scaling, no
data-path collective).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU sampler + CPU gather on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

N_NODES = 2_449_029
MEAN_DEG = 50.5
FEAT_DIM = 100
SIZES = [15, 10, 5]
BATCH = 1024
NCU_GATHER_TRAFFIC_BYTES = 726.9e6  # ncu --set full, bench batch (820 k rows x 400 B): 452.0 MB read + 274.9 MB written
WORKLOAD = "ogbn-products-shaped synthetic CSR (2449029 nodes, pareto(2) mean-deg 50.5), 1024 seeds, fanout [15,10,5], " \
           "100-d fp32 features"


def env_int(name, default):
    return int(os.environ.get(name, default))


# ----------------------------------------------------------------------------------------------------------------------
# synthetic workload (device-side generation; identical on every rank)
# ----------------------------------------------------------------------------------------------------------------------
def make_graph(device, n_nodes=N_NODES, mean_deg=MEAN_DEG, seed=0):
    g = torch.Generator(device=device).manual_seed(seed)
    raw = (1.0 - torch.rand(n_nodes, generator=g, device=device, dtype=torch.float64)).pow(-0.5)  # pareto(alpha=2)
    deg = (raw * (mean_deg / raw.mean())).floor().long().clamp_(max=n_nodes - 1)
    indptr = torch.zeros(n_nodes + 1, dtype=torch.long, device=device)
    indptr[1:] = deg.cumsum(0)
    n_edges = int(indptr[-1])
    row = torch.repeat_interleave(torch.arange(n_nodes, device=device), deg)
    col = torch.randint(0, n_nodes, (n_edges, ), generator=g, device=device)
    key, _ = torch.sort(row * n_nodes + col)  # columns sorted inside each row, as scipy's COO->CSR gives the reference
    indices = key % n_nodes
    return indptr, indices


def make_seed_batches(n_batches, n_nodes=N_NODES, batch=BATCH, seed=1):
    g = torch.Generator().manual_seed(seed)
    return [torch.randperm(n_nodes, generator=g)[:batch].pin_memory() if torch.cuda.is_available() else
            torch.randperm(n_nodes, generator=g)[:batch] for _ in range(n_batches)]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self.stop_flag = gpu_index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                               "-i", str(self.gpu_index)], text=True, timeout=5)
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU implementation of the path (oracle/_ref when built)
# ----------------------------------------------------------------------------------------------------------------------
def host_threads():
    """All the host threads this process may use.  torchrun exports OMP_NUM_THREADS=1, which would leave the reference's
    CPU gather (and its OpenMP build) on one core: undo that for the reference arm."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def reference_cpu_setup(indptr_cpu, indices_cpu):
    """The reference CPU extension compiled from its own sources (oracle/_ref).  Default: the as-shipped build (its
    at::parallel_for runs serially without -fopenmp, setup.py:58-69); QV_REF_VARIANT=omp selects the -fopenmp build (both
    register the same pybind types, so one process can only hold one; measured here: 0.40 s vs 0.46 s per batch for the
    sampler -- reindex_single is serial in both).  The CPU gather uses every host thread either way."""
    from oracle import oracle
    host_threads()
    order = (True, False) if os.environ.get("QV_REF_VARIANT", "") == "omp" else (False, True)
    for openmp in order:
        ext = oracle.load_reference(openmp=openmp)
        if ext is not None:
            return {"ext": ext, "openmp": openmp, "kind": "reference",
                    "quiver": ext.cpu_quiver_from_csr_array(indptr_cpu, indices_cpu)}
    # oracle/_ref was not built (it needs /root/reference at build time): fall back to the C restatement of the GPU path
    # ("port": same outputs as the product, one core) so that the baseline key is never empty
    return {"ext": None, "openmp": False, "kind": "port", "quiver": _OraclePort(oracle, indptr_cpu, indices_cpu)}


class _OraclePort:
    """sample_neighbor / reindex_single of oracle/qv_oracle.c behind the reference extension's call shapes."""

    def __init__(self, oracle, indptr, indices):
        self.o, self.indptr, self.indices = oracle, indptr.numpy(), indices.numpy()

    def sample_neighbor(self, nodes, k):
        out, cnt = self.o.sample_neighbor(self.indptr, self.indices, nodes.numpy(), int(k))
        return torch.from_numpy(out), torch.from_numpy(cnt)

    def reindex_single(self, nodes, out, cnt):
        return tuple(torch.from_numpy(a) for a in self.o.reindex(nodes.numpy(), out.numpy(), cnt.numpy()))


def reference_cpu_step(ref, seeds, x_cpu):
    """GraphSageSampler.sample restated over the reference's C++ bindings (sage_sampler.py:118-147, mode='CPU') +
    the CPU gather of bench_feature.py:62-66.  Returns (edges, rows, t_sample, t_gather)."""
    t0 = time.perf_counter()
    nodes, edges = seeds, 0
    for size in SIZES:
        out, cnt = ref["quiver"].sample_neighbor(nodes, size)
        frontier, row_idx, col_idx = ref["quiver"].reindex_single(nodes, out, cnt)
        edges += out.numel()
        nodes = frontier
    t1 = time.perf_counter()
    rows = x_cpu[nodes]
    t2 = time.perf_counter()
    return edges, rows.shape[0], t1 - t0, t2 - t1


def cpu_baseline_sample(indptr_cpu, indices_cpu, batches_host, warmup, x_cpu, row_bytes, n_b=4):
    """The `cpu_baseline` object of the N=1 line: the reference's CPU path timed on a bounded sample (n_b batches of the
    same workload, ~2 s of CPU work) on this box's host cores."""
    ref = reference_cpu_setup(indptr_cpu, indices_cpu)
    e = r = 0
    ts = tg = 0.0
    reference_cpu_step(ref, batches_host[0], x_cpu)
    t0 = time.perf_counter()
    for b in batches_host[warmup:warmup + n_b]:
        ee, rr, a, gg = reference_cpu_step(ref, b, x_cpu)
        e, r, ts, tg = e + ee, r + rr, ts + a, tg + gg
    tt = time.perf_counter() - t0
    what = ("reference CPU extension compiled from its sources (as shipped: serial at::parallel_for)"
            if ref["kind"] == "reference" else "oracle/qv_oracle.c port, 1 core (oracle/_ref not built)")
    return {"value": e / tt, "unit": "edges/s", "cores": torch.get_num_threads() if ref["openmp"] else 1,
            "kind": ref["kind"],
            "sample": f"{n_b} batches of the same workload; {what} for sample+reindex, torch CPU gather on "
                      f"{torch.get_num_threads()} threads; host has {os.cpu_count()} cores",
            "seps_sampler_only": e / ts, "feature_gather_GiBps": r * row_bytes / tg / 2**30}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    torch.manual_seed(0)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    indptr, indices = make_graph(dev)
    indptr_cpu, indices_cpu = indptr.cpu(), indices.cpu()
    del indptr, indices
    x_cpu = torch.rand(N_NODES, FEAT_DIM)
    batches = make_seed_batches(args.steps + args.warmup)
    ref = reference_cpu_setup(indptr_cpu, indices_cpu)
    if ref is None:
        return {"impl": "reference", "unavailable": "oracle/_ref (reference CPU extension) was not built"}
    for b in batches[:args.warmup]:
        reference_cpu_step(ref, b, x_cpu)
    edges = rows = 0
    ts = tg = 0.0
    t0 = time.perf_counter()
    for b in batches[args.warmup:]:
        e, r, a, g = reference_cpu_step(ref, b, x_cpu)
        edges, rows, ts, tg = edges + e, rows + r, ts + a, tg + g
    total = time.perf_counter() - t0
    cores = torch.get_num_threads() if ref["openmp"] else 1
    value = edges / total
    base = {"kind": ref["kind"], "cores": cores, "value": value, "unit": "edges/s",
            "sample": f"{args.steps} batches of the full workload, "
                      + ("reference CPU extension " if ref["kind"] == "reference" else "oracle/qv_oracle.c port ")
                      + ("(1 core)" if ref["kind"] == "port" else
                         f"({'-fopenmp' if ref['openmp'] else 'as shipped: at::parallel_for serial'})") + " + torch CPU gather "
                      f"({torch.get_num_threads()} threads); host has {os.cpu_count()} cores",
            "seps_sampler_only": edges / ts, "feature_gather_GiBps": rows * FEAT_DIM * 4 / tg / 2**30}
    return {"metric": "sampled_edges_per_s (k-hop sample + feature gather per step)", "value": value, "unit": "edges/s",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": {"workload": WORKLOAD, "where": "host CPU"},
            "cpu_baseline": base, "gpu_launches": 0,
            "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist

    import quiver
    import torch_quiver
    from torch_quiver import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    # ---- setup (untimed) ---------------------------------------------------------------------------------------------
    indptr, indices = make_graph(dev)
    n_edges = indices.numel()
    indptr_cpu, indices_cpu = indptr.cpu(), indices.cpu()
    del indptr, indices
    topo = quiver.CSRTopo(indptr=indptr_cpu, indices=indices_cpu)
    sampler = quiver.pyg.GraphSageSampler(topo, SIZES, device=local_rank, mode="GPU")
    sampler.overlap = args.overlap  # opt-in pipelining of sample(i+1) with gather(i) on a private stream (default off)
    sampler.inputs_ready = True  # the device-resident seed batches below are materialised before the timed region
    g = torch.Generator().manual_seed(7)
    x_cpu = torch.rand(N_NODES, FEAT_DIM, generator=g)
    if world == 1:
        feature = quiver.Feature(rank=local_rank, device_list=[local_rank], device_cache_size="2G",
                                 cache_policy="device_replicate", csr_topo=topo)  # 980 MB table: fully in HBM
        feature.from_cpu_tensor(x_cpu)
        placement = "1 GPU: whole table in local HBM, degree-ordered (feature_order folded into the gather)"
        remote_frac = 0.0
    else:
        quiver.init_p2p(list(range(world)))
        lo, hi = N_NODES * rank // world, N_NODES * (rank + 1) // world
        from quiver.shard_tensor import build_from_ranks
        store = build_from_ranks(x_cpu[lo:hi].contiguous(), local_rank)

        class _Sharded:  # Feature-shaped view over the rank-sharded store
            def __getitem__(self, idx):
                return store.gather(idx)
        feature = _Sharded()
        placement = f"{world}-way row shard over NVLink (CUDA IPC peer mappings), each rank gathers its own batch"
        remote_frac = 1.0 - 1.0 / world
    batches_host = make_seed_batches(args.steps + args.warmup, seed=1 + rank)
    batches_dev = [b.to(dev) for b in batches_host]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -----------------------------------------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    for b in batches_dev[:args.warmup]:
        n_id, _, adjs = sampler.sample(b)
        feature[n_id]
    barrier()
    # The timed regions below last ~10 ms -- shorter than one nvidia-smi poll.  Keep the SAME step loop running for
    # ~0.7 s first (untimed) so the clock / throttle record is taken under this workload's load.
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 0.7:
        for b in batches_dev[args.warmup:]:
            n_id, _, adjs = sampler.sample(b)
            feature[n_id]
    barrier()

    # ---- timed region A: K steps, inputs resident in HBM ("value"), with per-phase events ------------------------------
    want_overlap = sampler.overlap
    sampler.overlap = False
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    launches_s0 = _lib.launch_count()
    edges = rows = 0
    hop_bytes = 0  # B_hop = 40*E + 40*S + 8*F algorithmic bytes per hop
    nid_keep = []
    barrier()
    ev[0].record()
    for i, b in enumerate(batches_dev[args.warmup:]):
        n_id, _, adjs = sampler.sample(b)
        ev[3 * i + 1].record()
        res = feature[n_id]
        ev[3 * i + 2].record()
        edges += sum(a.edge_index.shape[1] for a in adjs)
        hop_bytes += sum(40 * a.edge_index.shape[1] + 40 * int(a.size[1]) + 8 * int(a.size[0]) for a in adjs)
        rows += n_id.numel()
        nid_keep.append(n_id)
        ev[3 * i + 3].record()
    barrier()
    launches_serial = _lib.launch_count() - launches_s0
    serial_ms = ev[0].elapsed_time(ev[3 * args.steps])
    sample_ms = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.steps))
    gather_ms = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.steps))
    sampler.overlap = want_overlap

    # ---- optional: the same K steps with the sampler on its private stream (--overlap) --------------------------------
    if want_overlap:
        launches0 = _lib.launch_count()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        edges_a = 0
        barrier()
        a0.record()
        for b in batches_dev[args.warmup:]:
            n_id, _, adjs = sampler.sample(b)
            res = feature[n_id]
            edges_a += sum(a.edge_index.shape[1] for a in adjs)
        a1.record()
        barrier()
        total_ms = a0.elapsed_time(a1)
        assert edges_a == edges  # same batches, same (deterministic) samples
        launches = _lib.launch_count() - launches0
    else:
        total_ms, launches = serial_ms, launches_serial

    # ---- timed region A': the same K steps through sample_and_gather (gather enqueued behind the last hop, frontier size
    #      read on the device: no GPU idle while the host learns the sizes).  Same batches, same results (asserted). ------
    fused_ms = None
    if not args.no_fuse and not want_overlap:
        fuse_target = feature if world == 1 else store
        for b in batches_dev[:args.warmup]:
            sampler.sample_and_gather(b, fuse_target)
        launches0 = _lib.launch_count()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        edges_f = rows_f = 0
        barrier()
        a0.record()
        for b in batches_dev[args.warmup:]:
            n_id, _, adjs, res = sampler.sample_and_gather(b, fuse_target)
            edges_f += sum(a.edge_index.shape[1] for a in adjs)
            rows_f += res.shape[0]
        a1.record()
        barrier()
        fused_ms = a0.elapsed_time(a1)
        assert edges_f == edges and rows_f == rows  # same batches, same (deterministic) samples
        assert torch.equal(res, feature[n_id])
        total_ms, launches = fused_ms, _lib.launch_count() - launches0

    # ---- timed region B: end to end through the public API with HOST seeds -------------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_edges = 0
    d2h = 0
    e0.record()
    for b in batches_host[args.warmup:]:
        n_id, _, adjs = sampler.sample(b)  # pinned host seeds -> H2D inside the call
        res = feature[n_id]
        probe = res[-1, :1].cpu()  # completes the step on the host (4 bytes) + the sampler's size read-back
        d2h = 4 + 8 * 4 * 9
        e2e_edges += sum(a.edge_index.shape[1] for a in adjs)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_fused_ms = 0.0
    if fused_ms is not None:  # informational: the same end-to-end loop through the fused extension call
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in batches_host[args.warmup:]:
            n_id, _, adjs, res = sampler.sample_and_gather(b, fuse_target)
            probe = res[-1, :1].cpu()
        e1.record()
        barrier()
        e2e_fused_ms = e0.elapsed_time(e1)
    clock_summary = clocks.summary()

    # ---- roofline of the dominant kernel (the gather): back-to-back launches over the timed batches' node lists -------
    row_bytes = FEAT_DIM * 4
    alg_bytes_per_row = 2 * row_bytes + 8 + (8 if world == 1 else 0)  # read + write + index (+ order)
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # (outputs are pre-allocated and the C-ABI call is issued directly so the host never starves the queue: the interval
    #  between the two events is back-to-back executions of the gather kernel and nothing else)
    st_raw = feature._my_store().shard_tensor if world == 1 else store.shard_tensor
    order = feature.feature_order if world == 1 else None
    outs = [torch.empty(n.numel(), FEAT_DIM, device=dev) for n in nid_keep[:2]]
    for j, n in enumerate(nid_keep[:2]):
        st_raw.gather(n, order, out=outs[j])
    barrier()
    reps = 0
    r0.record()
    for _ in range(3):
        for j, n in enumerate(nid_keep):
            st_raw.gather(n, order, out=outs[j % 2][:n.numel()] if n.numel() <= outs[j % 2].shape[0] else None)
            reps += 1
    r1.record()
    barrier()
    kern_ms = r0.elapsed_time(r1) / reps
    rows_per_launch = rows / args.steps
    achieved = rows_per_launch * alg_bytes_per_row / (kern_ms * 1e-3) / 1e9

    # ---- reduce over ranks -------------------------------------------------------------------------------------------
    stats = torch.tensor([total_ms, sample_ms, gather_ms, e2e_ms, kern_ms, serial_ms, e2e_fused_ms], dtype=torch.float64,
                         device=dev)
    # ---- a large-batch point (64 k seeds): the sampler where bandwidth, not launch latency, matters -------
    big = None
    if not args.no_large_batch:
        gbig = torch.Generator().manual_seed(99 + rank)
        big_batches = [torch.randperm(N_NODES, generator=gbig)[:65536].to(dev) for _ in range(3)]
        sampler.sample(big_batches[0])
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        big_edges = big_bytes = 0
        b0.record()
        for bb in big_batches:
            _, _, adjs = sampler.sample(bb)
            big_edges += sum(a.edge_index.shape[1] for a in adjs)
            big_bytes += sum(40 * a.edge_index.shape[1] + 40 * int(a.size[1]) + 8 * int(a.size[0]) for a in adjs)
        b1.record()
        barrier()
        big_ms = b0.elapsed_time(b1)
        big = {"seeds": 65536, "seps": big_edges / (big_ms * 1e-3), "algorithmic_GBps": big_bytes / (big_ms * 1e-3) / 1e9,
               "frac_of_hbm_peak": big_bytes / (big_ms * 1e-3) / 1e9 / hbm_peak, "ms_per_batch": big_ms / 3}
        del big_batches

    # ---- informational: the opt-in fast (non-reference-stream) sampler on the same batches ------------------------------
    fast = None
    if not args.no_large_batch:
        sampler.quiver.set_fast(True)
        for b in batches_dev[:2]:
            sampler.sample(b)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fe = 0
        f0.record()
        for b in batches_dev[args.warmup:]:
            _, _, adjs = sampler.sample(b)
            fe += sum(a.edge_index.shape[1] for a in adjs)
        f1.record()
        barrier()
        fast = {"seps_sampler_only": fe / (f0.elapsed_time(f1) * 1e-3), "sample_ms_per_step": f0.elapsed_time(f1) / args.steps,
                "note": "qv_sampler_set_fast: O(k) per row, NOT the reference's random stream; not part of `value`"}
        sampler.quiver.set_fast(False)

    sums = torch.tensor([edges, rows, e2e_edges, launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    total_ms, sample_ms, gather_ms, e2e_ms, kern_ms, serial_ms, e2e_fused_ms = stats.tolist()
    edges_all, rows_all, e2e_edges_all, launches_all = sums.tolist()
    if rank != 0:
        return None

    value = edges_all / (total_ms * 1e-3)
    out = {
        "metric": "sampled_edges_per_s (k-hop sample + feature gather per step)",
        "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_edges": n_edges, "placement": placement, "sampler_mode": "GPU (CSR in HBM), "
                   "reference-exact XORWOW sampling (rand_seed 0)",
                   "pipelining": ("sampler on its own high-priority stream (as the reference's stream pool): sample(i+1) "
                                  "overlaps the still-running feature gather of step i; every call returns completed "
                                  "results" if sampler.overlap else
                                  ("sample_and_gather (qv_khop_gather): the gather is enqueued behind the last hop with the "
                                   "frontier size read on the device; one stream, one host wait per step"
                                   if fused_ms is not None else "none: sampler and gather on one stream")), "l2": "inputs larger than L2 (990 MB CSR + 980 MB "
                   "feature table vs 126 MB L2); fresh seeds every step", "edges_per_step": edges_all / args.steps / world,
                   "rows_per_step": rows_all / args.steps / world},
        "seps_sampler_only": edges_all / (sample_ms * 1e-3),
        "feature_gather_GBps": rows_all * row_bytes / (gather_ms * 1e-3) / 1e9,
        "feature_gather_GiBps": rows_all * row_bytes / (gather_ms * 1e-3) / 2**30,
        "sample_ms_per_step": sample_ms / args.steps, "gather_ms_per_step": gather_ms / args.steps,
        "serial_ms_per_step": serial_ms / args.steps, "serial_edges_per_s": edges_all / (serial_ms * 1e-3),
        "sampler_roofline": {"bound": "hbm (nominally; at 1024 seeds the hops are launch/latency bound)",
                             "algorithmic_bytes_per_step": hop_bytes / args.steps, "formula": "sum over hops 40E+40S+8F",
                             "achieved": hop_bytes / (sample_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hop_bytes / (sample_ms * 1e-3) / 1e9 / hbm_peak, "large_batch": big},
        "fast_mode": fast,
        "gpu_launches": int(launches_all),
        "e2e": {"value": e2e_edges_all / (e2e_ms * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": BATCH * 8,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                "api": "sampler.sample(host seeds) then feature[n_id] -- the reference's two calls",
                "fused_value": (e2e_edges_all / (e2e_fused_ms * 1e-3)) if e2e_fused_ms > 0 else None},
        "clocks": clock_summary,
        "roofline": {"kernel": "gather_batch_flat_kernel<16,16> (feature gather, qv_gather.cu)", "bound": "hbm",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": NCU_GATHER_TRAFFIC_BYTES if world == 1 else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r1_gather_full.txt",
                     "peak_source": peak_src, "algorithmic_bytes_per_row": alg_bytes_per_row,
                     "rows_per_launch": rows_per_launch, "kernel_ms": kern_ms,
                     "how": "CUDA events around back-to-back launches of the timed batches' gathers (3 passes)"},
    }
    if world > 1:
        nv = rows_per_launch * remote_frac * row_bytes / (kern_ms * 1e-3) / 1e9
        out["nvlink"] = {"achieved_GBps_per_gpu_ingress": nv, "peak": 770.0, "frac": nv / 770.0,
                         "peak_source": "measured peer-copy 770 GB/s per direction (B200_PROFILING.md)",
                         "remote_row_fraction": remote_frac}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample(indptr_cpu, indices_cpu, batches_host, args.warmup, x_cpu, row_bytes)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-batch", action="store_true")
    ap.add_argument("--no-fuse", action="store_true",
                    help="`value` from sample() + feature[n_id] as two calls instead of sample_and_gather")
    ap.add_argument("--overlap", action="store_true",
                    help="run the sampler on its own high-priority stream so sample(i+1) overlaps gather(i)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner with printf on
    # the first communicator), so file descriptor 1 is pointed at stderr for the whole run and the result line goes to a
    # private duplicate of the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        res = run_reference(args, rank, world)
    else:
        if world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        res = run_ours(args, rank, world, local_rank)
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    if res is not None:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(res) + "\n").encode())
    os.close(real_stdout)


if __name__ == "__main__":
    main()
