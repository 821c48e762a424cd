"""Time a 2-hop sample whose seed list holds one 142 k-degree node (everything else tiny): the kernel time is that row's
generator chain.  Run with QV_MEGA=0 / QV_HEAVY_FIRST=0 to see what chain splitting and the longest-first schedule buy.
Usage: python profiles/mega_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import quiver  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    n = 300000
    deg = rng.integers(20, 60, n)
    deg[0] = 142000
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(1, n, int(indptr[-1])).astype(np.int64)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [5, 5], device=0, mode="GPU")
    seeds = torch.from_numpy(np.concatenate([[0], 1 + rng.permutation(n - 1)[:1023]])).cuda()
    for _ in range(5):
        sampler.sample(seeds)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        n_id, _, adjs = sampler.sample(seeds)
    e1.record()
    torch.cuda.synchronize()
    print(f"QV_MEGA={os.environ.get('QV_MEGA', '1')} QV_HEAVY_FIRST={os.environ.get('QV_HEAVY_FIRST', '1')}: "
          f"{e0.elapsed_time(e1) / 20 * 1000:.1f} us per 2-hop sample, {n_id.numel()} nodes")


if __name__ == "__main__":
    main()
