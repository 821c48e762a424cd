"""Sampler with the topology's `indices` in pinned host memory (mode="UVA") vs HBM (mode="GPU"), bench graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, quiver
from bench import SIZES, make_graph
from microbench import time_ms
indptr, indices = make_graph(torch.device("cuda"))
topo = quiver.CSRTopo(indptr=indptr.cpu(), indices=indices.cpu())
del indices
for mode in ("GPU", "UVA"):
    sampler = quiver.pyg.GraphSageSampler(topo, SIZES, device=0, mode=mode)
    seeds = [torch.randperm(topo.node_count, device="cuda")[:1024] for _ in range(4)]
    i = [0]
    def run():
        sampler.sample(seeds[i[0] % 4]); i[0] += 1
    ms = time_ms(run, reps=20, warm=3)
    n_id, _, adjs = sampler.sample(seeds[0])
    e = sum(a.edge_index.shape[1] for a in adjs)
    print(f"mode={mode}: {ms*1e3:.1f} us per sample(), {e/ms/1e3:.0f} M SEPS", flush=True)
    del sampler
