"""compute-sanitizer driver for the round-2 hop kernels (hop_sample_kernel incl. the heavy-row generator/tester ring with
named barriers, hop_reindex_kernel with its grid barriers, the fused gather): a small graph with two rows above kHeavyDeg,
3 fused k-hop calls (plain, with e_id, with the gather), each checked against the oracle.  Run as
    QV_COOP=1 compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/sanitize_khop.py
(QV_COOP=1: a cooperative launch, so that the instrumented kernels' co-residency is the runtime's promise, not ours)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")]

import torch_quiver as qv  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    rng = np.random.default_rng(7)
    n = 30000
    deg = rng.integers(0, 30, n)
    heavy = os.environ.get("SANITIZE_HEAVY", "1") != "0"  # 0: no row above kHeavyDeg (no generator/tester blocks)
    deg[11] = 9000 if heavy else 3000  # streamed through the ring (281 draws per lane)
    deg[222] = 4000 if heavy else 2500  # heavy, 125 draws per lane
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(0, n, int(indptr[-1]), dtype=np.int64)
    indices[::7] = 11  # the hub is a frequent neighbour: it is in every hop's frontier
    indices[3::11] = 222
    q = qv.device_quiver_from_csr_array(torch.from_numpy(indptr), torch.from_numpy(indices), torch.zeros(1, dtype=torch.long),
                                        0, True)
    table = torch.from_numpy(rng.integers(0, 100, (n, 64)).astype(np.float32))
    st = qv.ShardTensor(0)
    st.append(table, 0)
    sizes = [15, 10, 5]
    for it in range(3):
        seeds = rng.permutation(n)[:700]
        o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, sizes, with_eid=True)
        dev = torch.from_numpy(seeds).cuda()
        if it == 0:
            n_id, hops = q.sample_khop(dev, sizes)
            rows = None
        elif it == 1:
            n_id, hops = q.sample_khop(dev, sizes, with_eid=True)
            rows = None
        else:
            n_id, hops, rows = q.sample_khop(dev, sizes, gather=(st, None))
        torch.cuda.synchronize()
        assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid)), it
        for hop, (o_ei, o_size, o_pos) in zip(hops, o_adjs[::-1]):  # the oracle returns PyG's order (last hop first)
            assert torch.equal(hop[0].cpu(), torch.from_numpy(o_ei)), it
            if it == 1:
                assert torch.equal(hop[3].cpu(), torch.from_numpy(o_pos))
        if rows is not None:
            assert torch.equal(rows.cpu(), table[torch.from_numpy(o_nid)])
    print(f"sanitize_khop ok: 3 fused calls, {qv._lib.launch_count()} kernel launches, fused={os.environ.get('QV_KHOP_FUSED', '1')}, "
          f"heavy rows={os.environ.get('SANITIZE_HEAVY', '1')}")


if __name__ == "__main__":
    main()
