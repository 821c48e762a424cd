"""GPU parity on TWO devices (skipped on a single-GPU box): shards in peer HBM read one-sidedly over NVLink, the
p2p_clique_replicate policy, and the reference's DDP pattern -- Feature + sampler handed to one mp.spawn worker per GPU
(examples/multi_gpu/pyg/ogb-products/dist_sampling_ogb_products_quiver.py:85-99,158-163)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_peer_shards_both_directions(oracle):
    _need_two()
    import torch_quiver as qv
    assert qv.can_device_access_peer(0, 1) and qv.can_device_access_peer(1, 0)
    qv.init_p2p([0, 1])
    n, d = 40000, 100
    x = torch.from_numpy(np.random.default_rng(0).integers(0, 10, (n, d)).astype(np.float32))
    idx = torch.from_numpy(np.random.default_rng(1).integers(0, n, 60000))
    for home in (0, 1):
        st = qv.ShardTensor(home)
        st.append(x[:15000], 0)
        st.append(x[15000:30000], 1)
        st.append(x[30000:].clone(), -1)
        for variant in (1, 2):
            st.gather_variant = variant
            with torch.cuda.device(home):
                got = st[idx.to(f"cuda:{home}")]
            assert got.device.index == home and torch.equal(got.cpu(), x[idx])


@pytest.mark.parametrize("policy", ["p2p_clique_replicate", "device_replicate"])
def test_feature_two_gpu_policies(policy):
    _need_two()
    import quiver
    from graphs import powerlaw_csr
    n, d = 50000, 128
    indptr, indices = powerlaw_csr(n, 8.0, seed=2)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = torch.randn(n, d)
    quiver.init_p2p([0, 1])
    budget = 10000 * d * 4  # 10 k rows per GPU
    idx = torch.randint(0, n, (70000, ))
    for rank in (0, 1):
        topo.feature_order = None  # a Feature built from an already-ordered topo expects a pre-permuted tensor (feature.py:211-215)
        f = quiver.Feature(rank=rank, device_list=[0, 1], device_cache_size=budget, cache_policy=policy, csr_topo=topo)
        f.from_cpu_tensor(x)
        if policy == "p2p_clique_replicate":  # 20 k hot rows striped over the two GPUs, the rest on the host
            st = f.clique_tensor_list[0].shard_tensor
            assert st.device_count() == 3 and [s.rows for s in st.shards] == [10000, 10000, n - 20000]
        with torch.cuda.device(rank):
            res = f[idx.to(f"cuda:{rank}")]
        assert res.device.index == rank and torch.equal(res.cpu(), x[idx])


def _ddp_worker(rank, world, feature, sampler, x, seeds_all, want, ok):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "torch-quiver_b200")]
    torch.cuda.set_device(rank)
    seeds = seeds_all[rank]
    n_id, bs, adjs = sampler.sample(seeds)  # rebuilt lazily on this worker's GPU
    rows = feature[n_id]  # GPU shards arrive as CUDA IPC handles; peer rows are read over NVLink
    good = n_id.device.index == rank and rows.device.index == rank
    good = good and torch.equal(n_id.cpu(), want[rank]) and torch.equal(rows.cpu(), x[n_id.cpu()])
    ok[rank] = 1 if good else 0


def test_ddp_style_spawn_two_gpus():
    _need_two()
    import quiver
    import torch.multiprocessing as mp
    from graphs import powerlaw_csr
    n, d = 30000, 100
    indptr, indices = powerlaw_csr(n, 10.0, seed=3)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = torch.randn(n, d)
    quiver.init_p2p([0, 1])
    feature = quiver.Feature(rank=0, device_list=[0, 1], device_cache_size=8000 * d * 4,
                             cache_policy="p2p_clique_replicate", csr_topo=topo)
    feature.from_cpu_tensor(x)
    sampler = quiver.pyg.GraphSageSampler(topo, [10, 5], device=0, mode="GPU")
    seeds_all = [torch.arange(0, 512), torch.arange(1000, 1512)]
    want = [sampler.sample(s)[0].cpu() for s in seeds_all]
    ok = torch.zeros(2, dtype=torch.int32).share_memory_()
    mp.spawn(_ddp_worker, args=(2, feature, sampler, x, seeds_all, want, ok), nprocs=2, join=True)
    assert ok.tolist() == [1, 1]
