"""Generate the committed golden fixtures for the hot path FROM THE REFERENCE ITSELF (run in the build container).

  xorwow_kat.json   NVIDIA's curand_kernel.h XORWOW compiled for the host (oracle/_ref/curand_probe): the
                    third-party arithmetic behind the reference GPU sampler (cuda_random.cu.hpp:21-23,48).
  ref_cpu_kat.json  outputs of the reference's own CPU extension compiled unmodified from /root/reference
                    (oracle/_ref/torch_quiver_ref*.so; quiver.cpp:21-84): counts, verbatim rows (deg <= k),
                    reindex_single results, and one complete (unseeded) reference draw per graph that the structural
                    validator must accept.
  gpu_path_kat.json frozen outputs of the GPU-path restatement (oracle/qv_oracle.c: CSRRowWiseSampleKernel's generator
                    assignment + first-occurrence reindex, cuda_random.cu.hpp:7-69, quiver_sample.cu:18-63) for whole
                    k-hop samples, rand_seed 0.  NOT produced by the reference (its GPU build cannot run here and its
                    tests store no sampled ids): it freezes the oracle, which the two files above pin, so that neither a
                    kernel change nor an oracle change can move the sampled ids unnoticed.
Usage: python tests/golden/make_golden.py   (needs `make -C oracle all ref` first; /root/reference is NOT needed at
test time -- the tests read only the JSON files written here.)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from graphs import MINI, powerlaw_csr, simple_graph  # noqa: E402
from oracle import oracle  # noqa: E402


def xorwow():
    pairs = [(0, 0), (0, 1), (0, 31), (0, 32), (0, 127), (1, 0), (1, 33), (5, 64), (415, 127), (123456, 100),
             (2**33 + 7, 3), (17, 1000), (0, 2**20 + 5), (46875, 96), (2**32 - 1, 127)]
    res = oracle.curand_probe(16, pairs)
    assert res is not None, "build oracle/_ref/curand_probe first (make -C oracle)"
    json.dump({"source": "curand_kernel.h (CUDA 12.9) compiled for the host by oracle/curand_probe.cpp",
               "cases": res}, open(os.path.join(HERE, "xorwow_kat.json"), "w"), indent=0)
    print("xorwow_kat.json:", len(res), "cases")


def ref_cpu():
    import torch
    ref = oracle.load_reference()
    assert ref is not None, "build oracle/_ref first (python oracle/build_ref.py)"
    cases = []
    graphs = {
        "mini": (np.array(MINI["indptr"]), np.array(MINI["indices"])),
        "powerlaw_300": powerlaw_csr(300, 12.0, seed=3),
        "simple_100_10": simple_graph(100, 10),
    }
    rng = np.random.default_rng(11)
    for name, (indptr, indices) in graphs.items():
        n = indptr.shape[0] - 1
        cq = ref.cpu_quiver_from_csr_array(torch.from_numpy(indptr), torch.from_numpy(indices))
        for k in (2, 5, 25, n):
            seeds = np.array(MINI["seeds"]) if name == "mini" else rng.permutation(min(n, 300))[:64]
            out, cnt = cq.sample_neighbor(torch.from_numpy(seeds), int(k))
            frontier, row, col = cq.reindex_single(torch.from_numpy(seeds), out, cnt)
            cases.append(dict(graph=name, k=int(k), seeds=seeds.tolist(), counts=cnt.tolist(), draw=out.tolist(),
                              frontier=frontier.tolist(), row_idx=row.tolist(), col_idx=col.tolist()))
    json.dump({"source": "reference CPU extension (quiver.cpp:21-84) compiled unmodified from /root/reference",
               "graphs": {"powerlaw_300": dict(n_nodes=300, mean_deg=12.0, seed=3), "simple_100_10": dict(n=100, nbr=10)},
               "cases": cases}, open(os.path.join(HERE, "ref_cpu_kat.json"), "w"))
    print("ref_cpu_kat.json:", len(cases), "cases")


def gpu_path():
    cases = []
    for name, (n, mean, seed, S, sizes) in {
            "two_tiles": (500, 14.0, 21, 70, [5, 3]),          # 70 seeds: crosses the 64-row tile boundary
            "three_hops": (800, 20.0, 22, 40, [15, 10, 5]),    # the bench fan-out
            "wide_first_hop": (400, 30.0, 23, 33, [25, 10]),   # the Reddit fan-out (config 0)
    }.items():
        indptr, indices = powerlaw_csr(n, mean, seed=seed)
        seeds = np.random.default_rng(seed).permutation(n)[:S].astype(np.int64)
        n_id, bs, adjs = oracle.khop(indptr, indices, seeds, sizes)
        cases.append(dict(name=name, graph=dict(n_nodes=n, mean_deg=mean, seed=seed), seeds=seeds.tolist(), sizes=sizes,
                          n_id=n_id.tolist(),
                          adjs=[dict(edge_index=ei.tolist(), size=list(map(int, size))) for ei, size in adjs]))
    json.dump({"source": "oracle/qv_oracle.c GPU-path restatement, rand_seed 0 (see make_golden.py docstring)",
               "cases": cases}, open(os.path.join(HERE, "gpu_path_kat.json"), "w"))
    print("gpu_path_kat.json:", len(cases), "cases")


if __name__ == "__main__":
    xorwow()
    ref_cpu()
    gpu_path()
