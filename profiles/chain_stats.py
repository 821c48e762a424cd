"""How unbalanced is the reference's work decomposition?  For each hop of a bench batch: per virtual warp (rows 64b+w+4i)
the number of sequential reservoir iterations sum(ceil((deg-k)/32)), max vs mean."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch_quiver as qv
from bench import SIZES, make_graph
indptr, indices = make_graph(torch.device("cuda"))
deg_all = indptr[1:] - indptr[:-1]
print("max degree", int(deg_all.max()), "nodes>10k", int((deg_all > 10000).sum()), ">2k", int((deg_all > 2000).sum()))
q = qv.device_quiver_from_csr_array(indptr.cpu(), indices, None, 0, True)
seeds = torch.randperm(indptr.numel() - 1, device="cuda")[:1024]
nodes = seeds
for k in SIZES:
    deg = deg_all[nodes]
    it = torch.where(deg > k, (deg - k + 31) // 32, torch.zeros_like(deg))
    S = nodes.numel()
    pad = (-S) % 64
    itp = torch.cat([it, it.new_zeros(pad)]).view(-1, 16, 4)  # [block, i, w]
    chain = itp.sum(1).flatten()  # per virtual warp
    top = torch.topk(chain, 5).values.tolist()
    print(f"k={k} S={S} rows: total iters {int(it.sum())}, virtual warps {chain.numel()}, mean chain {chain.float().mean():.1f}, "
          f"max chains {top}, max single row {int(it.max())}")
    out, cnt = q.sample_neighbor(0, nodes, k)
    nodes, _, _ = q.reindex_single(nodes, out, cnt)
