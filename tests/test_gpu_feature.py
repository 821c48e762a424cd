"""GPU parity, feature gather: 0 ULP (pure byte copy) against the oracle and against `tensor[idx]`, which is the
reference's own assertion (tests/python/cuda/test_features.py:339,364,421; test_shard_tensor.py:77,106), across
dtypes, row sizes (16-byte clean and not), tiers, invalid ids, the folded feature_order, both kernel variants, and
the full quiver.Feature API incl. mp.spawn IPC."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(n, d, dtype, seed=0):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.integers(0, 10, (n, d)).astype(np.float32)).to(dtype)  # test_features.py:310-313


def _np(t):
    return t.view(torch.int16).numpy() if t.dtype in (torch.float16, torch.bfloat16) else t.numpy()


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("d", [100, 128, 256, 602, 768, 1, 3, 17, 600])
def test_single_shard_matches_oracle_and_indexing(oracle, d, dtype, variant):
    import torch_quiver as qv
    n, m = 20000, 30011
    x = _mk(n, d, dtype, seed=d)
    if variant == 2 and (d * x.element_size()) % 16 != 0:
        pytest.skip("TMA variant needs 16-byte rows")
    st = qv.ShardTensor(0)
    st.gather_variant = variant
    st.append(x, 0)
    idx = torch.from_numpy(np.random.default_rng(1).integers(0, n, m))
    got = st[idx.cuda()]
    assert got.shape == (m, d) and got.dtype == dtype and got.is_cuda and got.is_contiguous()
    assert torch.equal(got.cpu(), x[idx])
    want = oracle.gather([_np(x)], idx.numpy())
    assert np.array_equal(_np(got.cpu()), want)
    assert st.shape() == [n, d] and st.size(0) == n and st.size(1) == d and st.numel() == n * d
    assert st.stride(0) == d and st.device() == 0 and st.device_count() == 1


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("d,dtype", [(128, torch.float32), (602, torch.float32), (600, torch.float16), (100, torch.float32)])
def test_tiers_hbm_shards_plus_pinned_host(oracle, d, dtype, variant):
    import torch_quiver as qv
    if variant == 2 and (d * torch.empty(0, dtype=dtype).element_size()) % 16 != 0:
        pytest.skip("TMA variant needs 16-byte rows")
    n = 12000
    x = _mk(n, d, dtype, seed=3)
    cuts = [0, 5000, 5001, 9000, n]  # two real HBM shards, a 1-row shard, and the zero-copy host tier
    st = qv.ShardTensor(0)
    st.gather_variant = variant
    host_part = x[cuts[3]:].clone()
    for a, b in zip(cuts[:3], cuts[1:4]):
        st.append(x[a:b], 0)
    st.append(host_part, -1)
    assert st.device_count() == 4 and st.size(0) == n
    rng = np.random.default_rng(5)
    idx = np.concatenate([rng.integers(0, n, 9000), np.array(cuts[:-1]), np.array(cuts[1:]) - 1])
    got = st[torch.from_numpy(idx).cuda()]
    assert torch.equal(got.cpu(), x[idx])
    parts = [_np(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(_np(got.cpu()), oracle.gather(parts, idx))


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_invalid_ids_give_zero_rows_and_feature_order_is_folded(oracle, variant):
    import torch_quiver as qv
    n, d = 5000, 64
    x = _mk(n, d, torch.float32, seed=9) + 1.0  # no zero rows in the table
    st = qv.ShardTensor(0)
    st.gather_variant = variant
    st.append(x[:3000], 0)
    st.append(x[3000:].clone(), -1)
    idx = np.array([0, -1, n, n - 1, 2**40, -2**40, 2999, 3000], dtype=np.int64)
    got = st[torch.from_numpy(idx).cuda()].cpu()
    want = oracle.gather([x[:3000].numpy(), x[3000:].numpy()], idx)
    assert np.array_equal(got.numpy(), want)
    assert not got[[1, 2, 4, 5]].any() and torch.equal(got[[0, 3, 6, 7]], x[[0, n - 1, 2999, 3000]])
    order = torch.from_numpy(np.random.default_rng(0).permutation(n))
    ids = torch.from_numpy(np.random.default_rng(1).integers(0, n, 7777))
    got = st.gather(ids.cuda(), order.cuda())
    assert torch.equal(got.cpu(), x[order[ids]])
    assert np.array_equal(got.cpu().numpy(),
                          oracle.gather([x[:3000].numpy(), x[3000:].numpy()], ids.numpy(), order.numpy()))
    # empty request
    assert st[torch.empty(0, dtype=torch.long, device="cuda")].shape == (0, d)


def test_large_output_beyond_4gib_offsets():
    """The reference's unsigned 32-bit `warp_start * stride` wraps once the output exceeds 4 GiB
    (shard_tensor.cu.hpp:53); here every offset is 64-bit.  1.5 M rows x 3072 B = 4.6 GB."""
    import torch_quiver as qv
    n, d = 100000, 768
    x = torch.randn(n, d)
    st = qv.ShardTensor(0)
    st.append(x, 0)
    m = 1_500_000
    idx = torch.randint(0, n, (m, ), device="cuda")
    got = st[idx]
    xd = x.cuda()
    for lo in (0, m // 2, m - 4096):
        assert torch.equal(got[lo:lo + 4096], xd[idx[lo:lo + 4096]])
    assert torch.equal(got[-1], xd[idx[-1]])


@pytest.mark.parametrize("policy,cache", [("device_replicate", "2M"), ("p2p_clique_replicate", "1M"),
                                          ("device_replicate", 0), ("device_replicate", "1G")])
def test_feature_api_matches_tensor_indexing(policy, cache):
    import quiver
    from graphs import powerlaw_csr
    n, d = 30000, 100
    indptr, indices = powerlaw_csr(n, 10.0, seed=4)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = torch.randn(n, d)
    quiver.init_p2p([0])
    f = quiver.Feature(rank=0, device_list=[0], device_cache_size=cache, cache_policy=policy, csr_topo=topo)
    f.from_cpu_tensor(x)
    assert f.shape == [n, d] and f.size(0) == n and f.size(1) == d and f.dim() == 2
    idx = torch.randint(0, n, (80000, ))
    res = f[idx.cuda()]
    assert res.is_cuda and res.device.index == 0
    assert torch.equal(res.cpu(), x[idx])  # original ids in, original rows out: feature_order is hidden
    assert torch.equal(f[idx].cpu(), x[idx])  # CPU indices are moved to the rank, as the reference does
    xh = x.half()
    fh = quiver.Feature(rank=0, device_list=[0], device_cache_size=cache, cache_policy=policy)
    fh.from_cpu_tensor(xh)
    assert torch.equal(fh[idx.cuda()].cpu(), xh[idx])


def _ipc_child(rank, feature, sampler, x, indptr, indices, seeds, want_nid, ok):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "torch-quiver_b200")]
    torch.cuda.set_device(0)
    idx = torch.randint(0, x.shape[0], (20000, ), device="cuda")
    res = feature[idx]  # lazy_init_from_ipc_handle: opens the parent's shards through CUDA IPC
    good = torch.equal(res.cpu(), x[idx.cpu()])
    n_id, bs, adjs = sampler.sample(seeds)  # lazy_init_quiver in the child
    good = good and torch.equal(n_id.cpu(), want_nid) and bs == seeds.numel() and len(adjs) == 2
    ok[rank] = 1 if good else 0


def test_feature_and_sampler_cross_mp_spawn():
    """examples/multi_gpu/pyg/ogb-products/dist_sampling_ogb_products_quiver.py:158-163: Feature and sampler are
    handed to mp.spawn workers; GPU shards travel as CUDA IPC handles, the cold tier as shared memory."""
    import quiver
    import torch.multiprocessing as mp
    from graphs import powerlaw_csr
    n, d = 20000, 128
    indptr, indices = powerlaw_csr(n, 8.0, seed=6)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = torch.randn(n, d)
    feature = quiver.Feature(rank=0, device_list=[0], device_cache_size="4M", cache_policy="device_replicate",
                             csr_topo=topo)
    feature.from_cpu_tensor(x)
    sampler = quiver.pyg.GraphSageSampler(topo, [10, 5], device=0, mode="GPU")
    seeds = torch.arange(500, 756)
    want_nid, _, _ = sampler.sample(seeds)
    ok = torch.zeros(2, dtype=torch.int32).share_memory_()
    mp.spawn(_ipc_child, args=(feature, sampler, x, indptr, indices, seeds, want_nid.cpu(), ok), nprocs=2, join=True)
    assert ok.tolist() == [1, 1]


def test_shards_are_freed():
    import torch_quiver as qv
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(3):
        st = qv.ShardTensor(0)
        st.append(torch.zeros(200000, 256), 0)  # 205 MB
        st[torch.arange(10, device="cuda")]
        del st
    import gc
    gc.collect()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 100 * 2**20  # the reference never frees shards (no destructor): SURVEY.md 8(b)


def test_reddit_scale_sample_then_tiered_gather():
    """BASELINE config 0 shape (Reddit: 232,965 nodes, ~114 M edges, fan-out [25,10], 602-d fp32 = 2408-byte rows that are
    only 8-byte multiples) with the reference's bench placement "20 % of the rows cached on the GPU, the rest in host
    memory" (docs/Introduction_en.md:95-97): sampled n_id -> feature[n_id] must equal x[n_id] bit for bit, through the
    degree-ordered hot/cold split."""
    import quiver
    N, D = 232_965, 602
    g = torch.Generator(device="cuda").manual_seed(3)
    raw = (1.0 - torch.rand(N, generator=g, device="cuda", dtype=torch.float64)).pow(-0.5)
    deg = (raw * (491.5 / raw.mean())).floor().long().clamp_(max=N - 1)
    indptr = torch.zeros(N + 1, dtype=torch.long, device="cuda")
    indptr[1:] = deg.cumsum(0)
    indices = torch.randint(0, N, (int(indptr[-1]), ), generator=g, device="cuda")
    topo = quiver.CSRTopo(indptr=indptr.cpu(), indices=indices.cpu())
    del indices
    x = torch.rand(N, D)
    budget = int(0.2 * N) * D * 4
    feature = quiver.Feature(rank=0, device_list=[0], device_cache_size=budget, cache_policy="device_replicate",
                             csr_topo=topo)
    feature.from_cpu_tensor(x)
    st = feature.device_tensor_list[0].shard_tensor
    assert st.device_count() == 2 and st.size(0) == N  # hot HBM shard + zero-copy host tier
    sampler = quiver.pyg.GraphSageSampler(topo, [25, 10], device=0, mode="GPU")
    seeds = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:1024]
    n_id, bs, adjs = sampler.sample(seeds)
    assert bs == 1024 and torch.equal(n_id[:bs].cpu(), seeds)
    assert sum(a.edge_index.shape[1] for a in adjs) > 200_000
    rows = feature[n_id]
    assert rows.shape == (n_id.numel(), D)
    assert torch.equal(rows.cpu(), x[n_id.cpu()])


@pytest.mark.parametrize("policy,cache,d,dtype", [("device_replicate", "1G", 100, torch.float32),
                                                  ("device_replicate", "2M", 100, torch.float32),
                                                  ("device_replicate", 0, 128, torch.float16),
                                                  ("p2p_clique_replicate", "1M", 602, torch.float32)])
def test_sample_and_gather_equals_the_two_calls(oracle, policy, cache, d, dtype):
    """the fused sample -> gather (qv_khop_gather, frontier size read on the device) returns exactly
    what `sample(seeds)` followed by `feature[n_id]` returns -- ids vs the oracle k-hop, rows vs x[n_id], all tiers."""
    import quiver
    from graphs import powerlaw_csr
    n = 40000
    indptr, indices = powerlaw_csr(n, 12.0, seed=9)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = _mk(n, d, dtype, seed=5)
    f = quiver.Feature(rank=0, device_list=[0], device_cache_size=cache, cache_policy=policy, csr_topo=topo)
    f.from_cpu_tensor(x)
    sampler = quiver.pyg.GraphSageSampler(topo, [7, 5, 3], device=0, mode="GPU")
    for it, S in enumerate([1, 33, 1000, 257]):
        seeds = torch.randperm(n, generator=torch.Generator().manual_seed(it))[:S].cuda()
        n_id, bs, adjs, rows = sampler.sample_and_gather(seeds, f)
        n_id2, bs2, adjs2 = sampler.sample(seeds)
        assert bs == bs2 == S and torch.equal(n_id, n_id2) and len(adjs) == len(adjs2) == 3
        for a, b in zip(adjs, adjs2):
            assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.size, b.size)
        want_nid = oracle.khop(indptr, indices, seeds.cpu().numpy(), [7, 5, 3])[0]
        assert np.array_equal(n_id.cpu().numpy(), want_nid)
        assert rows.shape == (n_id.numel(), d) and rows.dtype == dtype and rows.is_contiguous()
        assert torch.equal(rows.cpu(), x[n_id.cpu()])
        assert torch.equal(rows, f[n_id])


def test_sample_and_gather_raw_shard_tensor_and_fallbacks():
    """A bare ShardTensor (no feature_order), the empty batch and the "-1 = all neighbours" fan-out (no static bound:
    the call falls back to the two separate calls) all give sample() + store[n_id]."""
    import quiver
    import torch_quiver as qv
    from graphs import powerlaw_csr
    n, d = 5000, 48
    indptr, indices = powerlaw_csr(n, 6.0, seed=2)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = _mk(n, d, torch.float32, seed=1)
    st = qv.ShardTensor(0)
    st.append(x[:3000], 0)
    st.append(x[3000:], -1)
    seeds = torch.arange(0, 600, 3).cuda()
    for sizes in ([4, 4], [3, -1]):
        sampler = quiver.pyg.GraphSageSampler(topo, sizes, device=0, mode="GPU")
        n_id, bs, adjs, rows = sampler.sample_and_gather(seeds, st)
        n_id2, _, adjs2 = sampler.sample(seeds)
        assert torch.equal(n_id, n_id2) and all(torch.equal(a.edge_index, b.edge_index) for a, b in zip(adjs, adjs2))
        assert torch.equal(rows.cpu(), x[n_id.cpu()])
    n_id, bs, adjs, rows = sampler.sample_and_gather(torch.empty(0, dtype=torch.long).cuda(), st)
    assert bs == 0 and n_id.numel() == 0 and rows.shape[0] == 0


def test_sample_and_gather_out_of_range_seed_takes_the_checked_path(oracle):
    """An id outside [0, N) makes the direct node map bail out; the call is redone on the hash path and the gather with
    it -- the result must still equal the per-call path (zero rows for the invalid id)."""
    import quiver
    from graphs import powerlaw_csr
    n, d = 3000, 20
    indptr, indices = powerlaw_csr(n, 5.0, seed=3)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = _mk(n, d, torch.float32, seed=2)
    f = quiver.Feature(rank=0, device_list=[0], device_cache_size="1G", cache_policy="device_replicate")
    f.from_cpu_tensor(x)
    sampler = quiver.pyg.GraphSageSampler(topo, [3, 3], device=0, mode="GPU")
    seeds = torch.tensor([5, n + 17, 9, 11]).cuda()
    n_id, bs, adjs, rows = sampler.sample_and_gather(seeds, f)
    n_id2, _, adjs2 = sampler.sample(seeds)
    assert torch.equal(n_id, n_id2)
    assert torch.equal(rows, f[n_id])
    assert torch.count_nonzero(rows[1]) == 0


def test_from_mmap_with_id_parts_and_saved_parts(tmp_path):
    """Feature.from_mmap (feature.py:95-192): parts given as row-id tensors into a numpy memmap, or as paths of saved row
    tensors (the `.pth` layout quiver_partition_feature writes, partition.py:234-247); set_local_order restores original
    ids (feature.py:283-294)."""
    import quiver
    from quiver.feature import DeviceConfig
    n, d = 20000, 64
    path = tmp_path / "feat.bin"
    arr = np.memmap(path, dtype=np.float32, mode="w+", shape=(n, d))
    arr[:] = np.random.default_rng(3).integers(0, 10, (n, d)).astype(np.float32)
    arr.flush()
    mm = np.memmap(path, dtype=np.float32, mode="r", shape=(n, d))
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0))
    hot, cold = perm[:6000], perm[6000:]
    idx = torch.randint(0, n, (50000, ))
    full = torch.from_numpy(np.array(mm))  # a writable copy
    want = full[idx]

    f = quiver.Feature(rank=0, device_list=[0], device_cache_size=0, cache_policy="device_replicate")
    f.from_mmap(mm, DeviceConfig([hot], cold))
    f.set_local_order(perm)
    assert f.shape == [n, d]
    assert torch.equal(f[idx.cuda()].cpu(), want)

    torch.save(full[hot].clone(), tmp_path / "gpu0.pth")
    torch.save(full[cold].clone(), tmp_path / "cpu.pth")
    g = quiver.Feature(rank=0, device_list=[0], device_cache_size=0, cache_policy="device_replicate")
    g.from_mmap(None, DeviceConfig([str(tmp_path / "gpu0.pth")], str(tmp_path / "cpu.pth")))
    g.set_local_order(perm)
    assert torch.equal(g[idx.cuda()].cpu(), want)


def test_shards_created_from_device_memory(oracle):
    """Extension for tables larger than host memory (C5 / north-star): a shard copied from a CUDA tensor (the reference
    refuses: CHECK_CPU, quiver_feature.cu:19,147) and a shard created empty in HBM and filled in place."""
    import torch_quiver as qv
    n, d = 50000, 100  # 400-byte rows: the shard pitch stays 400; d = 602 below pads 2408 -> 2416
    for d in (100, 602, 256):
        x = torch.from_numpy(np.random.default_rng(d).integers(0, 10, (n, d)).astype(np.float32))
        idx = torch.from_numpy(np.random.default_rng(1).integers(0, n, 70000)).cuda()
        st = qv.ShardTensor(0)
        st.append(x[:20000].cuda(), 0)                                   # device -> shard copy (pitched)
        view = st.append_empty(n - 20000 - 5000, [d], torch.float32, 0)  # in place: the caller writes the rows
        assert view.shape == (n - 25000, d) and view.is_cuda and view.stride(0) * 4 == st.shards[-1].pitch
        view.copy_(x[20000:n - 5000].cuda())
        st.append(x[n - 5000:].clone(), -1)                              # plus a pinned-host tier behind them
        assert st.shape() == [n, d] and st.device_count() == 3
        got = st[idx]
        assert torch.equal(got.cpu(), x[idx.cpu()])
        want = oracle.gather([x.numpy()], idx.cpu().numpy())
        assert np.array_equal(got.cpu().numpy(), want)
        items = st.share_ipc()  # both HBM shards are the library's own allocations: exportable
        assert len(items) == 2 and items[1].shape == [n - 25000, d]
    fp16 = qv.ShardTensor(0)
    v = fp16.append_empty(1000, [64], torch.float16, 0)
    ref = torch.randn(1000, 64).half()
    v.copy_(ref.cuda())
    sel = torch.arange(999, -1, -1).cuda()
    assert torch.equal(fp16[sel].cpu(), ref[sel.cpu()])
