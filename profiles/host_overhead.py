"""Host-side cost of the public API calls (tiny graph, so device work is launch-bound and the numbers are host time)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch, cProfile, pstats
import quiver
from graphs import powerlaw_csr
indptr, indices = powerlaw_csr(4000, 8.0, seed=1)
topo = quiver.CSRTopo(indptr=indptr, indices=indices)
sampler = quiver.pyg.GraphSageSampler(topo, [15, 10, 5], device=0, mode="GPU")
x = torch.randn(4000, 100)
feature = quiver.Feature(0, [0], "1G", csr_topo=topo); feature.from_cpu_tensor(x)
seeds = torch.arange(0, 64).cuda()
for _ in range(20):
    n_id, _, adjs = sampler.sample(seeds); feature[n_id]
torch.cuda.synchronize()
def loop(n):
    t0 = time.perf_counter()
    for _ in range(n):
        n_id, _, adjs = sampler.sample(seeds)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(n):
        r = feature[n_id]
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t1) / n * 1e6
print("sample() wall us, feature[] wall us:", loop(300))
q = sampler.quiver
t0 = time.perf_counter()
for _ in range(300):
    q.sample_khop(seeds, [15, 10, 5])
torch.cuda.synchronize(); print("sample_khop only us:", (time.perf_counter() - t0) / 300 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(300):
    n_id, _, adjs = sampler.sample(seeds); r = feature[n_id]
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
