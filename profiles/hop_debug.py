"""Device-side phase timing of the two per-hop kernels on a bench graph (QV_HOP_DEBUG=1 makes them printf).
Usage: QV_HOP_DEBUG=1 python profiles/hop_debug.py [config] [n_batches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import quiver
import bench
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "ns"]
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda")
indptr, indices = bench.make_graph(dev, cfg)
topo = quiver.CSRTopo(indptr=indptr, indices=indices)
sampler = quiver.pyg.GraphSageSampler(topo, cfg["sizes"], device=0, mode="GPU")
batches = [b.to(dev) for b in bench.make_seed_batches(nb + 2, cfg["n_nodes"], cfg["batch"], seed=1, legacy=cfg["legacy"])]
for b in batches:
    torch.cuda.synchronize()
    print("---- batch", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_id, _, adjs = sampler.sample(b)
    e1.record()
    torch.cuda.synchronize()
    print(f"sample: {e0.elapsed_time(e1) * 1e3:.1f} us, frontier {n_id.numel()}, edges {[a.edge_index.shape[1] for a in adjs]}", flush=True)
