"""Work decomposition of the reference-exact sampler on a bench graph (default: the north-star graph): per hop, the rows,
the draws, and how unbalanced the reference's per-warp generator chains are (rows 64b+w+4i share one warp's 32 streams).
Usage: python profiles/ns_chain_stats.py [config]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch_quiver as qv
import bench
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "ns"]
dev = torch.device("cuda")
indptr, indices = bench.make_graph(dev, cfg)
deg_all = indptr[1:] - indptr[:-1]
print("nodes", deg_all.numel(), "edges", indices.numel(), "max degree", int(deg_all.max()),
      "nodes>32k", int((deg_all > 32768).sum()), ">10k", int((deg_all > 10000).sum()), ">1536", int((deg_all > 1536).sum()))
q = qv.device_quiver_from_csr_array(indptr, indices, None, 0, True)
for trial in range(2):
    nodes = bench.make_seed_batches(1, cfg["n_nodes"], cfg["batch"], seed=5 + trial)[0].to(dev)
    for k in cfg["sizes"]:
        deg = deg_all[nodes]
        it = torch.where(deg > k, (deg - k + 31) // 32, torch.zeros_like(deg))
        S = nodes.numel()
        pad = (-S) % 64
        itp = torch.cat([it, it.new_zeros(pad)]).view(-1, 16, 4)  # [block, i, w]
        chain = itp.sum(1).flatten()  # per virtual warp
        top = torch.topk(chain, 5).values.tolist()
        hist = [int((deg > t).sum()) for t in (k, 64, 256, 1536, 8192, 32768)]
        print(f"k={k} S={S}: draw rounds total {int(it.sum())} (mean/row {it.float().mean():.2f}), warps {chain.numel()}, mean chain "
              f"{chain.float().mean():.1f}, top chains {top}, max row {int(it.max())}; rows with deg > (k,64,256,1536,8192,32768): {hist}")
        out, cnt = q.sample_neighbor(0, nodes, k)
        nodes, _, _ = q.reindex_single(nodes, out, cnt)
    print("frontier", nodes.numel())
