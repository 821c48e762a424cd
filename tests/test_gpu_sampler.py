"""GPU parity, sampler: the CUDA path (through the C ABI) must equal the oracle BIT FOR BIT -- counts, sampled ids
under the reference's generator seed, reindex outputs, the fused k-hop -- plus structural / property checks at
BASELINE.json sizes where the CPU oracle would take too long."""
import numpy as np
import pytest
import torch

from graphs import MINI, powerlaw_csr, simple_graph

pytestmark = pytest.mark.gpu


def _quiver(indptr, indices, cuda=True, device=0):
    import torch_quiver as qv
    return qv.device_quiver_from_csr_array(torch.from_numpy(indptr), torch.from_numpy(indices),
                                           torch.zeros(1, dtype=torch.long), device, cuda)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).cuda()


@pytest.fixture(scope="module")
def g2k():
    return powerlaw_csr(2000, 30.0, seed=7)


@pytest.mark.parametrize("k", [1, 2, 5, 25, 33, 64, 2000, -1])
@pytest.mark.parametrize("S", [1, 63, 64, 65, 1000])
def test_sample_neighbor_bit_exact(oracle, g2k, k, S):
    indptr, indices = g2k
    q = _quiver(indptr, indices)
    seeds = np.random.default_rng(S * 131 + k).integers(0, 2000, S)
    out, cnt = q.sample_neighbor(0, _dev(seeds), k)
    ref_out, ref_cnt = oracle.sample_neighbor(indptr, indices, seeds, k)
    assert cnt.cpu().numpy().tolist() == ref_cnt.tolist()
    assert out.cpu().numpy().tolist() == ref_out.tolist()
    assert out.dtype == torch.int64 and out.is_cuda and cnt.shape == (S, )


def test_sample_heavy_tail_and_large_fanout(oracle):
    # hub rows far above k (long reservoir loops), fan-outs beyond the shared-memory slot limit (1024)
    indptr, indices = powerlaw_csr(6000, 120.0, seed=8, alpha=1.3)
    assert np.diff(indptr).max() > 3000
    q = _quiver(indptr, indices)
    seeds = np.argsort(-np.diff(indptr))[:200].copy()  # the 200 biggest hubs
    for k in (10, 1024, 1025, 1500):
        out, cnt = q.sample_neighbor(0, _dev(seeds), k)
        ref_out, ref_cnt = oracle.sample_neighbor(indptr, indices, seeds, k)
        assert torch.equal(cnt.cpu(), torch.from_numpy(ref_cnt))
        assert torch.equal(out.cpu(), torch.from_numpy(ref_out)), k
        assert oracle.validate_sample(indptr, indices, seeds, k, ref_cnt, out.cpu().numpy()) == 0


def test_many_blocks_and_other_seeds(oracle):
    # > 1024 blocks of 64 seeds: grows the XORWOW state cache; rand_seed != 0 uses per-launch states
    indptr, indices = powerlaw_csr(50000, 12.0, seed=9)
    q = _quiver(indptr, indices)
    seeds = np.random.default_rng(3).integers(0, 50000, 70000)
    out, cnt = q.sample_neighbor(0, _dev(seeds), 4)
    ref_out, ref_cnt = oracle.sample_neighbor(indptr, indices, seeds, 4)
    assert torch.equal(out.cpu(), torch.from_numpy(ref_out)) and torch.equal(cnt.cpu(), torch.from_numpy(ref_cnt))
    small = seeds[:3000]
    for rs in (1, 12345, 2**40 + 3):
        q.rand_seed = rs
        out, _ = q.sample_neighbor(0, _dev(small), 4)
        ref_out, _ = oracle.sample_neighbor(indptr, indices, small, 4, rand_seed=rs)
        assert torch.equal(out.cpu(), torch.from_numpy(ref_out)), rs
    q.rand_seed = 0
    out2, _ = q.sample_neighbor(0, _dev(small), 4)
    ref2, _ = oracle.sample_neighbor(indptr, indices, small, 4)
    assert torch.equal(out2.cpu(), torch.from_numpy(ref2))


def test_edge_cases(oracle, g2k):
    indptr, indices = g2k
    q = _quiver(indptr, indices)
    # empty seed list (the reference launches a 0-block grid: quiver.cu.hpp:388)
    out, cnt = q.sample_neighbor(0, torch.empty(0, dtype=torch.long, device="cuda"), 5)
    assert out.numel() == 0 and cnt.numel() == 0
    f, r, c = q.reindex_single(torch.empty(0, dtype=torch.long, device="cuda"), out, cnt)
    assert f.numel() == 0 and r.numel() == 0 and c.numel() == 0
    # isolated seeds only
    iso = np.nonzero(np.diff(indptr) == 0)[0][:10]
    assert len(iso) > 0
    out, cnt = q.sample_neighbor(0, _dev(iso), 5)
    assert out.numel() == 0 and cnt.sum().item() == 0
    f, r, c = q.reindex_single(_dev(iso), out, cnt)
    assert f.cpu().tolist() == iso.tolist()
    # k = 0, out-of-range seeds (defined here as degree 0), wrong dtype / device
    out, cnt = q.sample_neighbor(0, _dev([1, 2, 3]), 0)
    assert out.numel() == 0
    out, cnt = q.sample_neighbor(0, _dev([-5, 2000, 10**12, 3]), 5)
    assert cnt.cpu().tolist()[:3] == [0, 0, 0]
    with pytest.raises(RuntimeError):
        q.sample_neighbor(0, torch.tensor([1, 2], dtype=torch.int32, device="cuda"), 2)
    with pytest.raises(RuntimeError):
        q.sample_neighbor(0, torch.tensor([1, 2]), 2)


def test_mini_known_answer_on_gpu():
    m = MINI
    q = _quiver(np.array(m["indptr"]), np.array(m["indices"]))
    out, cnt = q.sample_neighbor(0, _dev(m["seeds"]), m["k"])
    assert cnt.cpu().tolist() == m["counts"]
    assert out.cpu().tolist()[4:6] == [0, 2]  # seed 1 has deg 2 <= k: verbatim CSR row
    f, r, c = q.reindex_single(_dev(m["seeds"]), _dev(m["draw"]), _dev(m["counts"]))
    assert (f.cpu().tolist(), r.cpu().tolist(), c.cpu().tolist()) == (m["frontier"], m["row_idx"], m["col_idx"])


@pytest.mark.parametrize("n,nbr,k", [(10, 5, 10), (100, 10, 5), (1000, 10, 10)])  # tests/cpp/test_quiver_cpu.cpp:70-75
def test_reference_structural_cases_on_gpu(oracle, n, nbr, k):
    indptr, indices = simple_graph(n, nbr)
    q = _quiver(indptr, indices)
    seeds = np.arange(n)
    out, cnt = q.sample_neighbor(0, _dev(seeds), k)
    assert oracle.validate_sample(indptr, indices, seeds, k, cnt.cpu().numpy(), out.cpu().numpy()) == 0


def test_reindex_bit_exact(oracle, golden_dir):
    import json
    import os
    rng = np.random.default_rng(4)
    q = _quiver(*powerlaw_csr(100, 3.0, seed=1))
    # reference CPU extension's own outputs (unique seeds)
    kat = json.load(open(os.path.join(golden_dir, "ref_cpu_kat.json")))
    for c in kat["cases"]:
        f, r, col = q.reindex_single(_dev(c["seeds"]), _dev(c["draw"]), _dev(c["counts"]))
        assert f.cpu().tolist() == c["frontier"] and r.cpu().tolist() == c["row_idx"]
        assert col.cpu().tolist() == c["col_idx"]
    # random, heavy duplication, duplicate seeds (GPU semantics: merged), sizes across scan-tile boundaries
    for S, tot_per in [(1, 3), (1023, 1), (1024, 2), (1025, 7), (5000, 25), (40000, 10)]:
        seeds = rng.integers(0, max(S // 2, 2), S)
        counts = rng.integers(0, tot_per + 1, S)
        outputs = rng.integers(0, max(S, 50), int(counts.sum()))
        f, r, col = q.reindex_single(_dev(seeds), _dev(outputs), _dev(counts))
        of, orow, ocol = oracle.reindex(seeds, outputs, counts)
        assert torch.equal(f.cpu(), torch.from_numpy(of)), S
        assert torch.equal(r.cpu(), torch.from_numpy(orow)), S
        assert torch.equal(col.cpu(), torch.from_numpy(ocol)), S


@pytest.mark.parametrize("reindex", ["map", "hash"])
@pytest.mark.parametrize("sizes", [[25, 10], [15, 10, 5], [3], [2, 2, 2, 2], [40, 0, 3]])
def test_fused_khop_equals_oracle_and_per_hop(oracle, sizes, reindex, monkeypatch):
    import quiver
    monkeypatch.setenv("QV_KHOP_REINDEX", reindex)  # direct node map (default) vs per-hop hash table
    indptr, indices = powerlaw_csr(20000, 25.0, seed=10)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, sizes, device=0, mode="GPU")
    seeds = np.random.default_rng(len(sizes)).permutation(20000)[:512]
    n_id, bs, adjs = sampler.sample(torch.from_numpy(seeds))
    o_nid, o_bs, o_adjs = oracle.khop(indptr, indices, seeds, sizes)
    assert bs == o_bs == 512
    assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
    assert len(adjs) == len(sizes)
    for adj, (o_ei, o_size) in zip(adjs, o_adjs):
        assert adj.edge_index.is_contiguous() and adj.edge_index.shape == (2, o_ei.shape[1])
        assert torch.equal(adj.edge_index.cpu(), torch.from_numpy(o_ei))
        assert adj.size.tolist() == list(o_size) and not adj.size.is_cuda
        assert adj.e_id.numel() == 0
    assert torch.equal(n_id[:bs].cpu(), torch.from_numpy(seeds))  # targets come first
    # the per-hop path (what the reference's Python loop does) gives the same answer
    sampler.fused = False
    n_id2, _, adjs2 = sampler.sample(torch.from_numpy(seeds))
    assert torch.equal(n_id2, n_id)
    for a, b in zip(adjs, adjs2):
        assert torch.equal(a.edge_index, b.edge_index) and a.size.tolist() == b.size.tolist()


def test_fused_khop_repeated_calls_and_bad_seeds(oracle):
    """The node map persists across calls (reset after every sample); invalid seeds reroute the call to the hash path."""
    import quiver
    indptr, indices = powerlaw_csr(8000, 15.0, seed=14)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [6, 4], device=0, mode="GPU")
    rng = np.random.default_rng(0)
    for it in range(4):
        seeds = rng.permutation(8000)[:300] if it != 2 else np.concatenate([rng.integers(0, 8000, 200)] * 2)  # dups
        n_id, bs, adjs = sampler.sample(torch.from_numpy(seeds))
        o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, [6, 4])
        assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid)), it
        for adj, (o_ei, _) in zip(adjs, o_adjs):
            assert torch.equal(adj.edge_index.cpu(), torch.from_numpy(o_ei)), it
    bad = torch.tensor([5, 8000, 17, -3, 42, 2**40])
    n_id, bs, adjs = sampler.sample(bad)
    sampler.fused = False
    n_id2, _, adjs2 = sampler.sample(bad)
    assert torch.equal(n_id, n_id2) and all(torch.equal(a.edge_index, b.edge_index) for a, b in zip(adjs, adjs2))
    sampler.fused = True
    seeds = rng.permutation(8000)[:300]  # the map must be clean again after the rerouted call
    n_id, _, _ = sampler.sample(torch.from_numpy(seeds))
    assert torch.equal(n_id.cpu(), torch.from_numpy(oracle.khop(indptr, indices, seeds, [6, 4])[0]))


def test_full_neighbourhood_hop_falls_back(oracle):
    import quiver
    indptr, indices = powerlaw_csr(3000, 10.0, seed=12)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [5, -1], device=0, mode="GPU")
    seeds = np.arange(100, 228)
    n_id, bs, adjs = sampler.sample(torch.from_numpy(seeds))
    o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, [5, 3000])
    assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
    assert torch.equal(adjs[0].edge_index.cpu(), torch.from_numpy(o_adjs[0][0]))


def test_uva_mode_reads_host_indices(oracle, g2k):
    indptr, indices = g2k
    q = _quiver(indptr, indices, cuda=False)  # indices stay in (registered) host memory
    seeds = np.random.default_rng(0).integers(0, 2000, 500)
    out, cnt = q.sample_neighbor(0, _dev(seeds), 10)
    ref_out, ref_cnt = oracle.sample_neighbor(indptr, indices, seeds, 10)
    assert torch.equal(out.cpu(), torch.from_numpy(ref_out)) and torch.equal(cnt.cpu(), torch.from_numpy(ref_cnt))


def test_non_default_stream_is_respected(oracle, g2k):
    indptr, indices = g2k
    q = _quiver(indptr, indices)
    seeds = np.random.default_rng(1).integers(0, 2000, 4096)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        d = _dev(seeds)
        out, cnt = q.sample_neighbor(0, d, 7)
        f, r, c = q.reindex_single(d, out, cnt)
    s.synchronize()
    ref_out, ref_cnt = oracle.sample_neighbor(indptr, indices, seeds, 7)
    of, _, ocol = oracle.reindex(seeds, ref_out, ref_cnt)
    assert torch.equal(out.cpu(), torch.from_numpy(ref_out))
    assert torch.equal(f.cpu(), torch.from_numpy(of)) and torch.equal(c.cpu(), torch.from_numpy(ocol))


def test_cal_neighbor_prob(oracle):
    indptr, indices = powerlaw_csr(5000, 9.0, seed=13)
    q = _quiver(indptr, indices)
    p = np.random.default_rng(2).random(5000).astype(np.float32)
    cur = torch.zeros(5000, device="cuda")
    q.cal_neighbor_prob(0, torch.from_numpy(p).cuda(), cur, 4)
    want = oracle.cal_next(p, 4, indptr, indices)
    # fp32, neighbours multiplied in CSR order, the final 1 - (1-p)*acc as one FMA (as nvcc contracts it in the
    # reference kernel): bit-identical
    assert np.array_equal(cur.cpu().numpy(), want)
    # two hops through the Python API (GraphSageSampler.sample_prob, sage_sampler.py:149-157)
    import quiver
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [4, 3], device=0, mode="GPU")
    train_idx = torch.arange(0, 5000, 7)
    prob = sampler.sample_prob(train_idx, 5000)
    p0 = np.zeros(5000, np.float32)
    p0[train_idx.numpy()] = 1
    want2 = oracle.cal_next(oracle.cal_next(p0, 4, indptr, indices), 3, indptr, indices)
    assert np.array_equal(prob.cpu().numpy(), want2)


def test_products_scale_properties():
    """BASELINE config 1 (ogbn-products-shaped: 2.45 M nodes, ~124 M edges, fan-out [15,10,5]) -- size-independent
    properties checked on the device: counts == min(deg,k); every sampled id is a neighbour of its seed; no position is
    picked twice; frontier is duplicate-free, starts with the seeds, and edge_index indexes inside it."""
    import quiver
    N = 2_449_029
    g = torch.Generator(device="cuda").manual_seed(0)
    raw = (1.0 - torch.rand(N, generator=g, device="cuda", dtype=torch.float64)).pow(-0.5)
    deg = (raw * (50.5 / raw.mean())).floor().long().clamp_(max=N - 1)
    indptr = torch.zeros(N + 1, dtype=torch.long, device="cuda")
    indptr[1:] = deg.cumsum(0)
    E = int(indptr[-1])
    row = torch.repeat_interleave(torch.arange(N, device="cuda"), deg)
    col = torch.randint(0, N, (E, ), generator=g, device="cuda")
    key, _ = torch.sort(row * N + col)  # CSR with sorted columns, as one flat sorted key array
    indices = key % N
    del row, col
    topo = quiver.CSRTopo(indptr=indptr.cpu(), indices=indices.cpu())
    sampler = quiver.pyg.GraphSageSampler(topo, [15, 10, 5], device=0, mode="GPU")
    seeds = torch.randperm(N, generator=g, device="cuda")[:1024]
    n_id, bs, adjs = sampler.sample(seeds)
    assert torch.equal(n_id[:bs], seeds)
    assert torch.unique(n_id).numel() == n_id.numel()
    n_dst = bs
    for adj, k in zip(adjs[::-1], [15, 10, 5]):
        src, dst = adj.edge_index[0], adj.edge_index[1]
        assert adj.size.tolist()[1] == n_dst
        assert int(src.max()) < adj.size[0] and int(dst.max()) < n_dst
        cnt = torch.bincount(dst, minlength=n_dst)
        assert torch.equal(cnt, deg[n_id[:n_dst]].clamp(max=k))
        assert bool((dst[1:] >= dst[:-1]).all())  # per-seed blocks are contiguous and in seed order
        # membership: (target node, source node) must be an edge
        ekey = n_id[dst] * N + n_id[src]
        pos = torch.searchsorted(key, ekey)
        assert bool((key[pos.clamp(max=E - 1)] == ekey).all())
        n_dst = int(adj.size[0])
    assert n_dst == n_id.numel()


@pytest.mark.parametrize("k", [31, 32, 33])
def test_small_kernel_boundary_and_multigraph(oracle, k):
    """Fan-outs around the 32-wide kernel switch, on a multigraph (duplicate neighbour ids inside a row, unsorted rows)."""
    rng = np.random.default_rng(21)
    n = 3000
    deg = rng.integers(0, 200, n)
    deg[:5] = [0, 1, 31, 32, 33]
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(0, 40, int(indptr[-1])).astype(np.int64)  # ids from a tiny range: heavy duplication
    q = _quiver(indptr, indices)
    seeds = np.concatenate([np.arange(5), rng.integers(0, n, 700)])
    out, cnt = q.sample_neighbor(0, _dev(seeds), k)
    ref_out, ref_cnt = oracle.sample_neighbor(indptr, indices, seeds, k)
    assert torch.equal(cnt.cpu(), torch.from_numpy(ref_cnt)) and torch.equal(out.cpu(), torch.from_numpy(ref_out))
    assert oracle.validate_sample(indptr, indices, seeds, k, ref_cnt, out.cpu().numpy()) == 0
    f, r, c = q.sample_sub(0, _dev(seeds), k)  # Quiver.sample_sub == sample_neighbor + reindex_single
    of, orow, ocol = oracle.reindex(seeds, ref_out, ref_cnt)
    assert torch.equal(f.cpu(), torch.from_numpy(of)) and torch.equal(r.cpu(), torch.from_numpy(orow))
    assert torch.equal(c.cpu(), torch.from_numpy(ocol))


def test_fused_khop_in_uva_mode(oracle):
    import quiver
    indptr, indices = powerlaw_csr(15000, 20.0, seed=15)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [8, 4, 2], device=0, mode="UVA")  # indices read zero-copy from host
    seeds = np.random.default_rng(5).permutation(15000)[:400]
    n_id, bs, adjs = sampler.sample(torch.from_numpy(seeds))
    o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, [8, 4, 2])
    assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
    for adj, (o_ei, _) in zip(adjs, o_adjs):
        assert torch.equal(adj.edge_index.cpu(), torch.from_numpy(o_ei))


def test_overlapped_sampling_gives_identical_results(oracle):
    import quiver
    indptr, indices = powerlaw_csr(12000, 18.0, seed=16)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [10, 5], device=0, mode="GPU")
    sampler.overlap = True  # private high-priority stream
    x = torch.randn(12000, 64)
    feature = quiver.Feature(0, [0], device_cache_size="1G")
    feature.from_cpu_tensor(x)
    rng = np.random.default_rng(6)
    keep = []
    for it in range(5):
        seeds = rng.permutation(12000)[:512]
        n_id, _, adjs = sampler.sample(torch.from_numpy(seeds) if it % 2 else torch.from_numpy(seeds).cuda())
        keep.append((seeds, n_id, adjs, feature[n_id]))  # the gather runs while the next sample() is issued
    for seeds, n_id, adjs, rows in keep:
        o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, [10, 5])
        assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
        assert all(torch.equal(a.edge_index.cpu(), torch.from_numpy(o[0])) for a, o in zip(adjs, o_adjs))
        assert torch.equal(rows.cpu(), x[torch.from_numpy(o_nid)])


def test_fast_mode_is_valid_uniform_and_not_the_reference_stream(oracle):
    """Opt-in O(k)-per-row sampling (Quiver.set_fast): NOT bit-compatible with the reference -- checked at parity level L2
    (SURVEY.md 8(c)): counts, membership, no duplicate positions, verbatim rows when deg <= k, uniformity, and the
    reindex / k-hop structure around it."""
    import quiver
    indptr, indices = powerlaw_csr(20000, 40.0, seed=17, alpha=1.5)
    q = _quiver(indptr, indices)
    q.set_fast(True)
    seeds = np.random.default_rng(8).permutation(20000)[:3000]
    for k in (1, 5, 25, 40, 200, -1):
        out, cnt = q.sample_neighbor(0, _dev(seeds), k)
        ref_cnt, _, tot = oracle.sample_counts(indptr, seeds, k)
        assert torch.equal(cnt.cpu(), torch.from_numpy(ref_cnt)) and out.numel() == tot
        assert oracle.validate_sample(indptr, indices, seeds, k, ref_cnt, out.cpu().numpy()) == 0, k
    out2, _ = q.sample_neighbor(0, _dev(seeds), 5)
    out3, _ = q.sample_neighbor(0, _dev(seeds), 5)
    assert not torch.equal(out2, out3)  # every call draws a fresh sample
    ref5, _ = oracle.sample_neighbor(indptr, indices, seeds, 5)
    assert not torch.equal(out2.cpu(), torch.from_numpy(ref5))
    # uniformity: one row of degree 97 with distinct neighbour ids, k = 6, many calls
    ip = np.array([0, 97], dtype=np.int64)
    idx = np.arange(1000, 1097, dtype=np.int64)
    q1 = _quiver(ip, idx)
    q1.set_fast(True)
    hits = np.zeros(97)
    row = _dev(np.zeros(512, dtype=np.int64))  # 512 copies of the row per call: same node, different calls differ
    trials = 0
    for _ in range(40):
        o, c = q1.sample_neighbor(0, row[:1], 6)
        picks = o.cpu().numpy()
        assert len(set(picks.tolist())) == 6
        hits[picks - 1000] += 1
        trials += 1
    for _ in range(960):
        o, _ = q1.sample_neighbor(0, row[:1], 6)
        hits[o.cpu().numpy() - 1000] += 1
        trials += 1
    expect = trials * 6 / 97
    chi2 = ((hits - expect) ** 2 / expect).sum()
    assert chi2 < 170.0, chi2  # 96 dof: p(chi2 > 170) ~ 5e-6
    # fused k-hop in fast mode: structure must be self-consistent
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [8, 4], device=0, mode="GPU")
    sampler.quiver.set_fast(True)
    s_t = torch.from_numpy(seeds[:500])
    n_id, bs, adjs = sampler.sample(s_t)
    assert torch.equal(n_id[:bs].cpu(), s_t) and torch.unique(n_id).numel() == n_id.numel()
    key = torch.from_numpy(np.repeat(np.arange(20000), np.diff(indptr)) * 20000 + indices).cuda().sort().values
    n_dst = bs
    deg = torch.from_numpy(np.diff(indptr)).cuda()
    for adj, k in zip(adjs[::-1], [8, 4]):
        src, dst = adj.edge_index
        assert torch.equal(torch.bincount(dst, minlength=n_dst), deg[n_id[:n_dst]].clamp(max=k))
        ekey = n_id[dst] * 20000 + n_id[src]
        pos = torch.searchsorted(key, ekey).clamp(max=key.numel() - 1)
        assert bool((key[pos] == ekey).all())
        n_dst = int(adj.size[0])
    assert n_dst == n_id.numel()


def _compact_rows(indptr, indices, rows):
    """CSR of just `rows` (row i of the result = rows[i]): sampling depends on a row's position in the seed list, its
    degree and its contents -- not on where the row lives -- so the oracle can be run on this instead of a multi-GB CSR."""
    start = indptr[rows]
    deg = indptr[rows + 1] - start
    cptr = torch.zeros(rows.numel() + 1, dtype=torch.long, device=rows.device)
    cptr[1:] = deg.cumsum(0)
    tot = int(cptr[-1])
    src = torch.repeat_interleave(start - cptr[:-1], deg) + torch.arange(tot, device=rows.device)
    return cptr.cpu().numpy(), indices[src].cpu().numpy()


def test_more_than_2_31_edges_bit_exact(oracle):
    """Maximum sizes: a CSR with 2.2e9 edges (17.6 GB of int64 ids, papers100M-symmetrised scale).  Every row offset past
    edge 2^31 needs 64-bit arithmetic end to end (the reference's thrust path carries int offsets in places).
    Sampled ids are compared bit for bit with the oracle run on the compacted rows of the seeds; the fused k-hop is
    compared with the per-hop calls and checked for membership."""
    import torch_quiver as qv
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2**30:
        pytest.skip("needs ~40 GB of free HBM")
    N = 4_000_000
    deg = torch.where(torch.arange(N, device="cuda") % 2 == 0, 5, 1095)  # copy path and reservoir path, mean 550
    indptr = torch.zeros(N + 1, dtype=torch.long, device="cuda")
    indptr[1:] = deg.cumsum(0)
    E = int(indptr[-1])
    assert E > 2**31
    indices = torch.empty(E, dtype=torch.long, device="cuda")
    step = 1 << 27
    for lo in range(0, E, step):  # ids from a formula, filled in place chunk by chunk
        e = torch.arange(lo, min(lo + step, E), device="cuda")
        indices[lo:lo + step] = (e * 2654435761 + (e >> 9)) % N
    del e
    q = qv.device_quiver_from_csr_array(indptr, indices, None, 0, True)
    rows = torch.cat([N - 1 - torch.arange(0, 3000, device="cuda") * 7,          # offsets around 2.2e9
                      int(N * 0.977) + torch.arange(0, 900, device="cuda"),  # just past 2^31
                      torch.arange(0, 500, device="cuda") * 11])                   # and the low end
    assert int(indptr[rows].max()) > 2**31
    cptr, cidx = _compact_rows(indptr, indices, rows)
    local = np.arange(rows.numel(), dtype=np.int64)
    for k in (3, 15, 40):
        out, cnt = q.sample_neighbor(0, rows, k)
        o_out, o_cnt = oracle.sample_neighbor(cptr, cidx, local, k)
        assert np.array_equal(cnt.cpu().numpy(), o_cnt) and np.array_equal(out.cpu().numpy(), o_out), k
    # fused k-hop == per-hop calls, and every edge is an edge of the graph
    nodes, hops_ref = rows, []
    for k in (15, 10):
        out, cnt = q.sample_neighbor(0, nodes, k)
        frontier, row_idx, col_idx = q.reindex_single(nodes, out, cnt)
        hops_ref.append(torch.stack([col_idx, row_idx]))
        nodes = frontier
    n_id, hops = q.sample_khop(rows, [15, 10])
    assert torch.equal(n_id, nodes)
    for (ei, n_src, n_dst), want in zip(hops, hops_ref):
        assert torch.equal(ei, want)
        tgt, srcn = n_id[ei[1]], n_id[ei[0]]
        # srcn must sit in tgt's row: indices[p] == srcn for some p in [indptr[tgt], indptr[tgt+1]) -- invert the formula
        # row by row on a sample of edges
        pick = torch.randperm(ei.shape[1], device="cuda")[:2000]
        for t, s in zip(tgt[pick].tolist(), srcn[pick].tolist()):
            lo, hi = int(indptr[t]), int(indptr[t + 1])
            assert bool((indices[lo:hi] == s).any())


def test_rmat_graph_khop_and_gather_bit_exact(oracle):
    """BASELINE config 2 shape at 1/16 scale (R-MAT (0.57,0.19,0.19,0.05), 0.6 M nodes, ~8 M edges with duplicate
    neighbours and ~half the rows empty, fan-out [15,10,5], 256-d fp32): k-hop ids and edge_index bit-exact with the
    oracle, gathered rows equal x[n_id]."""
    import quiver
    from graphs import rmat_csr
    n = 600_000
    indptr, indices = rmat_csr(20, 10_000_000, n_nodes=n, seed=2)
    assert (np.diff(indptr) == 0).mean() > 0.2 and np.diff(indptr).max() > 5_000
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [15, 10, 5], device=0, mode="GPU")
    x = torch.from_numpy(np.random.default_rng(0).integers(0, 10, (n, 256)).astype(np.float32))
    feature = quiver.Feature(rank=0, device_list=[0], device_cache_size="300M", cache_policy="device_replicate",
                             csr_topo=topo)  # ~half of the rows in HBM (degree-ordered), the rest in pinned host memory
    feature.from_cpu_tensor(x)
    rng = np.random.default_rng(5)
    for it in range(2):
        seeds = rng.permutation(n)[:1024]
        n_id, bs, adjs, rows = sampler.sample_and_gather(torch.from_numpy(seeds), feature)
        o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, [15, 10, 5])
        assert bs == 1024 and torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
        for adj, (o_ei, o_size) in zip(adjs, o_adjs):
            assert torch.equal(adj.edge_index.cpu(), torch.from_numpy(o_ei)) and adj.size.tolist() == list(o_size)
        assert torch.equal(rows.cpu(), x[n_id.cpu()])


@pytest.mark.parametrize("n_heavy", [3, 40, 300])
def test_longest_first_schedule_keeps_results(oracle, n_heavy):
    """The fused k-hop runs warps that own a heavy row (> 48 draws per lane) in extra blocks at the front of the grid
    (count_scan lists them, the regular slot retires).  Results must not move: few heavy rows, several in one warp
    (rows r and r+4 share a warp), and more than the 64-entry list can hold (the overflow stays on the regular
    schedule) -- ids and edge_index bit-exact with the oracle."""
    import quiver
    rng = np.random.default_rng(n_heavy)
    n = 30000
    deg = rng.integers(0, 12, n)
    deg[:n_heavy] = rng.integers(1700, 4000, n_heavy)   # heavy: (deg - k) / 32 > 48
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(0, n, int(indptr[-1])).astype(np.int64)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [5, 3], device=0, mode="GPU")
    # heavy nodes packed at the front of the seed list: rows 0, 4, 8, ... share warp 0 of tile 0
    seeds = np.concatenate([np.arange(n_heavy), n_heavy + rng.permutation(n - n_heavy)[:700]]).astype(np.int64)
    for _ in range(2):
        n_id, bs, adjs = sampler.sample(torch.from_numpy(seeds))
        o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, [5, 3])
        assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
        for adj, (o_ei, o_size) in zip(adjs, o_adjs):
            assert torch.equal(adj.edge_index.cpu(), torch.from_numpy(o_ei)) and adj.size.tolist() == list(o_size)
        seeds = seeds[::-1].copy()  # heavy rows at the END of the list the second time (last tiles of the grid)


@pytest.mark.parametrize("n_mega,sizes", [(1, [5, 3]), (3, [15, 10]), (11, [5, 3])])
def test_mega_rows_chain_splitting_is_bit_exact(oracle, n_mega, sizes):
    """Rows with more than 1024 draws per lane (degree > 32 k) are not walked by their warp: front-of-grid workers run
    256-draw segments from generator states obtained by GF(2) jump-ahead, the owner jumps over the row.  Must equal the
    sequential chain exactly: one mega row, several (two in one warp: seed positions 0 and 4; one as a warp's last row:
    position 60), more than the 8 the launch can list, and mega rows again in the second hop (the frontier keeps the seeds
    in front).  Degrees up to 150 k."""
    import quiver
    rng = np.random.default_rng(100 + n_mega)
    n = 60000
    deg = rng.integers(0, 10, n)
    mega_deg = rng.integers(33_000, 60_000, n_mega)
    mega_deg[0] = 150_000
    deg[:n_mega] = mega_deg
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(0, n, int(indptr[-1])).astype(np.int64)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, sizes, device=0, mode="GPU")
    others = n_mega + rng.permutation(n - n_mega)[:400]
    seeds = others.copy()
    slots = [0, 4, 60, 7, 129, 130, 200, 201, 202, 203, 300][:n_mega]  # seed positions of the mega nodes
    seeds[slots] = np.arange(n_mega)
    seeds = seeds.astype(np.int64)
    assert len(set(seeds.tolist())) == len(seeds)
    for _ in range(2):
        n_id, bs, adjs = sampler.sample(torch.from_numpy(seeds))
        o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, sizes)
        assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
        for adj, (o_ei, o_size) in zip(adjs, o_adjs):
            assert torch.equal(adj.edge_index.cpu(), torch.from_numpy(o_ei)) and adj.size.tolist() == list(o_size)


def test_khop_against_committed_golden_vectors(golden_dir):
    """GPU k-hop vs tests/golden/gpu_path_kat.json -- stored vectors, no oracle involved at run time."""
    import json
    import os
    import quiver
    kat = json.load(open(os.path.join(golden_dir, "gpu_path_kat.json")))
    for c in kat["cases"]:
        g = c["graph"]
        indptr, indices = powerlaw_csr(g["n_nodes"], g["mean_deg"], seed=g["seed"])
        topo = quiver.CSRTopo(indptr=indptr, indices=indices)
        for fused in (True, False):
            sampler = quiver.pyg.GraphSageSampler(topo, c["sizes"], device=0, mode="GPU")
            sampler.fused = fused
            n_id, bs, adjs = sampler.sample(torch.tensor(c["seeds"]))
            assert n_id.cpu().tolist() == c["n_id"] and bs == len(c["seeds"]), (c["name"], fused)
            for adj, want in zip(adjs, c["adjs"]):
                assert adj.edge_index.cpu().tolist() == want["edge_index"] and adj.size.tolist() == want["size"]


def test_pinned_host_seeds_are_read_in_place(oracle):
    """Seeds in pinned host memory go to the fused call as they are (hop 0's kernels read them over PCIe): same results as
    device seeds and as unpinned host seeds (which are copied), through sample() and sample_and_gather()."""
    import quiver
    indptr, indices = powerlaw_csr(20000, 25.0, seed=10)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    sampler = quiver.pyg.GraphSageSampler(topo, [15, 10, 5], device=0, mode="GPU")
    x = torch.from_numpy(np.random.default_rng(0).integers(0, 10, (20000, 16)).astype(np.float32))
    feature = quiver.Feature(rank=0, device_list=[0], device_cache_size="64M", csr_topo=topo)
    feature.from_cpu_tensor(x)
    seeds = torch.from_numpy(np.random.default_rng(3).permutation(20000)[:700])
    want_nid, _, want_adjs = sampler.sample(seeds.cuda())
    o_nid, _, _ = oracle.khop(indptr, indices, seeds.numpy(), [15, 10, 5])
    assert torch.equal(want_nid.cpu(), torch.from_numpy(o_nid))
    for s in (seeds.pin_memory(), seeds):
        n_id, bs, adjs = sampler.sample(s)
        assert bs == 700 and torch.equal(n_id, want_nid)
        assert all(torch.equal(a.edge_index, b.edge_index) for a, b in zip(adjs, want_adjs))
        n_id2, _, adjs2, rows = sampler.sample_and_gather(s, feature)
        assert torch.equal(n_id2, want_nid) and torch.equal(rows.cpu(), x[want_nid.cpu()])


@pytest.mark.timeout(180)
def test_two_samplers_on_one_device_sample_concurrently(oracle):
    """The per-hop reindex kernel synchronises its whole grid; two of them in flight on one device (two samplers, two
    threads, two streams) must not starve each other of SMs: with a second sampler alive on the device the kernel is
    launched cooperatively.  Results stay bit-exact."""
    import threading
    import quiver
    indptr, indices = powerlaw_csr(60000, 30.0, seed=12)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    samplers = [quiver.pyg.GraphSageSampler(topo, [15, 10, 5], device=0, mode="GPU") for _ in range(2)]
    seeds = [np.random.default_rng(40 + t).permutation(60000)[:1024] for t in range(2)]
    want = [oracle.khop(indptr, indices, s, [15, 10, 5])[0] for s in seeds]
    errors = []

    def work(t):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                dev_seeds = torch.from_numpy(seeds[t]).cuda()
                for _ in range(60):
                    n_id, _, _ = samplers[t].sample(dev_seeds)
                if not torch.equal(n_id.cpu(), torch.from_numpy(want[t])):
                    errors.append(f"thread {t}: wrong n_id")
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(t, )) for t in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=150)
    assert not any(th.is_alive() for th in threads), "the two samplers dead-locked"
    assert not errors, errors


@pytest.mark.parametrize("rand_seed", [1, 2**33 + 5])
def test_fused_khop_with_another_generator_seed(oracle, rand_seed):
    """rand_seed != 0 (the reference hard-codes 0): the per-launch generator states depend on each hop's row count, which
    only the device knows in the fused call -- same ids as the oracle's hop loop with that seed, heavy rows included."""
    import torch_quiver as qv
    indptr, indices = powerlaw_csr(30000, 40.0, seed=15, alpha=1.4)
    assert np.diff(indptr).max() > 3072  # above the heavy-list threshold: streamed rows too
    q = qv.device_quiver_from_csr_array(torch.from_numpy(indptr), torch.from_numpy(indices), None, 0, True)
    q.rand_seed = rand_seed
    hubs = np.argsort(-np.diff(indptr))[:5]
    rest = np.random.default_rng(9).permutation(30000)[:600]
    seeds = np.concatenate([hubs, rest[~np.isin(rest, hubs)]])
    n_id, hops = q.sample_khop(_dev(seeds), [10, 5, 3])
    o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, [10, 5, 3], rand_seed=rand_seed)
    assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
    for (ei, _, _), (o_ei, _) in zip(hops, o_adjs[::-1]):
        assert torch.equal(ei.cpu(), torch.from_numpy(o_ei))


def test_heavy_list_grows_past_its_initial_cap(oracle):
    """Fused k-hop, hop_sample_kernel: rows above 3072 neighbours are served by front-of-grid heavy blocks, 256 of them on a
    sampler's first call; a call that sees more (R-MAT frontiers do) leaves the overflow on the regular schedule and the
    sampler widens the front for the next call (up to 2048).  700 such rows, several per logical warp, heavy rows also
    first met on hop 1: ids, edge_index and e_id bit-exact with the oracle on the overflowing call and on the grown ones."""
    rng = np.random.default_rng(77)
    n = 40000
    n_heavy = 700
    deg = rng.integers(0, 10, n)
    deg[:n_heavy] = rng.integers(3100, 5000, n_heavy)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(0, n, int(indptr[-1])).astype(np.int64)
    indices[rng.random(indices.shape[0]) < 0.3] = rng.integers(0, n_heavy, 1)[0]  # one heavy hub met again and again
    light = indptr[n_heavy]
    indices[light::5] = rng.integers(0, n_heavy, indices[light::5].shape[0])  # light rows lead to heavy ones on hop 1
    q = _quiver(indptr, indices)
    sizes = [6, 4, 3]
    for it in range(3):
        seeds = np.concatenate([rng.permutation(n_heavy)[:300], n_heavy + rng.permutation(n - n_heavy)[:500]]).astype(np.int64)
        n_id, hops = q.sample_khop(_dev(seeds), sizes, with_eid=True)
        o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, sizes, with_eid=True)
        assert torch.equal(n_id.cpu(), torch.from_numpy(o_nid)), it
        for hop, (o_ei, _, o_pos) in zip(hops, o_adjs[::-1]):
            assert torch.equal(hop[0].cpu(), torch.from_numpy(o_ei)), it
            assert torch.equal(hop[3].cpu(), torch.from_numpy(o_pos)), it


def test_compiled_call_path_equals_the_ctypes_path(oracle, monkeypatch):
    """Quiver.sample_khop runs its host side compiled (csrc/pybind: khop_raw) when the adapter is built -- which build()
    does, so on a GPU box it must be there.  Same C calls underneath: ids, edge_index, e_id and gathered rows identical to
    the ctypes path and to the oracle; unsupported requests and bad seeds end in the same exceptions."""
    import torch_quiver as qv
    assert qv._compiled is not None, "the compiled call path is not built (python torch-quiver_b200/csrc/pybind/build.py)"
    indptr, indices = powerlaw_csr(20000, 30.0, seed=31)
    q = _quiver(indptr, indices)
    table = torch.from_numpy(np.random.default_rng(0).integers(0, 99, (20000, 24)).astype(np.float32))
    st = qv.ShardTensor(0)
    st.append(table, 0)
    order = torch.from_numpy(np.random.default_rng(1).permutation(20000)).cuda()
    seeds = np.random.default_rng(2).permutation(20000)[:777]
    pinned = torch.from_numpy(seeds).pin_memory()
    sizes = [7, 5, 3]
    o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, sizes, with_eid=True)

    def run():
        return (q.sample_khop(_dev(seeds), sizes), q.sample_khop(pinned, sizes, with_eid=True),
                q.sample_khop(_dev(seeds), sizes, gather=(st, order)), q.sample_khop(pinned, sizes, gather=(st, None), with_eid=True))

    fast = run()
    monkeypatch.setattr(qv, "_compiled", None)
    slow = run()
    for f, s in zip(fast, slow):
        assert torch.equal(f[0], s[0]) and torch.equal(f[0].cpu(), torch.from_numpy(o_nid))
        assert len(f) == len(s) and len(f[1]) == len(s[1]) == 3
        for hf, hs, (o_ei, _, o_pos) in zip(f[1], s[1], o_adjs[::-1]):
            assert len(hf) == len(hs) and torch.equal(hf[0], hs[0]) and tuple(hf[1:3]) == tuple(hs[1:3])
            assert torch.equal(hf[0].cpu(), torch.from_numpy(o_ei))
            if len(hf) > 3:
                assert torch.equal(hf[3], hs[3]) and torch.equal(hf[3].cpu(), torch.from_numpy(o_pos))
        if len(f) > 2:
            assert torch.equal(f[2], s[2])
    assert torch.equal(fast[2][2].cpu(), table[order.cpu()[torch.from_numpy(o_nid)]])
    assert torch.equal(fast[3][2].cpu(), table[torch.from_numpy(o_nid)])
    monkeypatch.undo()
    for compiled in (True, False):
        if not compiled:
            monkeypatch.setattr(qv, "_compiled", None)
        with pytest.raises(qv.Unsupported):
            q.sample_khop(_dev(seeds), [5, -1])
        with pytest.raises(RuntimeError):
            q.sample_khop(torch.from_numpy(seeds), sizes)  # pageable host memory is not device-visible
