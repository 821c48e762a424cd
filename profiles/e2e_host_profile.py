"""Where does the end-to-end step spend its host time?  North-star graph at 1/10 of the nodes (same per-step work shape),
pinned host seeds, sample_and_gather + 4-byte read-back; wall-clock split of one step, averaged over 200 steps."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import quiver
import bench
from quiver.shard_tensor import build_tiered_inplace
import torch_quiver as qv
from torch_quiver import _lib

cfg = dict(bench.CONFIGS["ns"], n_nodes=10_000_000)
dev = torch.device("cuda", 0)
indptr, indices = bench.make_graph(dev, cfg)
topo = quiver.CSRTopo(indptr=indptr, indices=indices)
sampler = quiver.pyg.GraphSageSampler(topo, cfg["sizes"], device=0, mode="GPU")
n, dim = cfg["n_nodes"], 256
store, _ = build_tiered_inplace(0, n, [dim], torch.float32, lambda v, lo, hi: v.fill_(1.0))
feature = quiver.Feature.from_tiered_store(0, store, None)
batches = bench.make_seed_batches(64, n, 1024, seed=1)
probe = torch.empty(1).pin_memory()
for b in batches[:8]:
    sampler.sample_and_gather(b, feature)
torch.cuda.synchronize()

# (1) whole step
t0 = time.perf_counter()
for _ in range(4):
    for b in batches[:50]:
        n_id, _, adjs, res = sampler.sample_and_gather(b, feature)
        probe.copy_(res[-1, :1], non_blocking=True)
        torch.cuda.current_stream().synchronize()
whole = (time.perf_counter() - t0) / 200
# (2) the C call alone, everything pre-allocated (what a compiled adapter with cached buffers would cost)
q = sampler.quiver
orig = _lib.lib.qv_khop_gather
t_c = [0.0]
def timed(*a):
    s = time.perf_counter(); r = orig(*a); t_c[0] += time.perf_counter() - s; return r
_lib.lib.qv_khop_gather = timed
for _ in range(4):
    for b in batches[:50]:
        n_id, _, adjs, res = sampler.sample_and_gather(b, feature)
        probe.copy_(res[-1, :1], non_blocking=True)
        torch.cuda.current_stream().synchronize()
_lib.lib.qv_khop_gather = orig
c_call = t_c[0] / 200
# (3) device time of the same step (events, no host read-back)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
bd = [b.cuda() for b in batches[:50]]
torch.cuda.synchronize(); e0.record()
for _ in range(4):
    for b in bd:
        sampler.sample_and_gather(b, feature)
e1.record(); torch.cuda.synchronize()
devt = e0.elapsed_time(e1) / 200 * 1e-3
# (4) python around the C call
t0 = time.perf_counter()
for _ in range(4):
    for b in batches[:50]:
        res[-1, :1]
py_slice = (time.perf_counter() - t0) / 200
print(f"whole e2e step {whole * 1e6:.1f} us | inside qv_khop_gather (launch + wait for the sampler's sizes) {c_call * 1e6:.1f} us | "
      f"device-timed step (no read-back) {devt * 1e6:.1f} us | python outside the C call {1e6 * (whole - c_call):.1f} us "
      f"(of which the tail waits for the gather)")
