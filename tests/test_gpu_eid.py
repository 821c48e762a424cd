"""GPU parity, SURVEY 8(f-3): the opt-in edge-id (`e_id`) output and `sample_sub` as one C call.

The reference plumbs edge ids through its sampler (quiver.cu.hpp:90-126, quiver_sample.cu:434-453) and then returns an
empty e_id (sage_sampler.py:143).  Here e_id[e] = CSR position of sampled edge e (or edge_ids[position]), checked against
the oracle's restatement of the same walk (qo_sample_neighbor_gpu_pos) bit for bit, on every sampling kernel: the
small-fan-out kernel (k <= 32, incl. heavy and mega rows), the generic kernel (k > 32, k > 1024), the fused k-hop, and
the opt-in fast sampler (structural)."""
import numpy as np
import pytest
import torch

from graphs import powerlaw_csr

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).cuda()


def _quiver(indptr, indices, edge_ids=None, cuda=True):
    import torch_quiver as qv
    eid = torch.zeros(1, dtype=torch.long) if edge_ids is None else torch.from_numpy(edge_ids)
    return qv.device_quiver_from_csr_array(torch.from_numpy(indptr), torch.from_numpy(indices), eid, 0, cuda)


@pytest.mark.parametrize("k", [1, 5, 25, 32, 33, 200, 1500])
def test_sample_neighbor_eid_is_the_csr_position(oracle, k):
    indptr, indices = powerlaw_csr(6000, 60.0, seed=21, alpha=1.4)
    q = _quiver(indptr, indices)
    seeds = np.concatenate([np.argsort(-np.diff(indptr))[:100], np.random.default_rng(k).integers(0, 6000, 900)])
    out, cnt, eid = q.sample_neighbor(0, _dev(seeds), k, return_eid=True)
    r_out, r_cnt, r_pos = oracle.sample_neighbor_pos(indptr, indices, seeds, k)
    assert torch.equal(out.cpu(), torch.from_numpy(r_out)) and torch.equal(cnt.cpu(), torch.from_numpy(r_cnt))
    assert torch.equal(eid.cpu(), torch.from_numpy(r_pos))
    assert np.array_equal(indices[eid.cpu().numpy()], out.cpu().numpy())  # the position holds the sampled neighbour
    # without the flag the call is unchanged
    out2, cnt2 = q.sample_neighbor(0, _dev(seeds), k)
    assert torch.equal(out2, out) and torch.equal(cnt2, cnt)


def test_user_edge_ids_are_mapped(oracle):
    indptr, indices = powerlaw_csr(3000, 20.0, seed=22)
    user = np.random.default_rng(0).permutation(indices.shape[0]).astype(np.int64) + 10**12
    seeds = np.random.default_rng(1).integers(0, 3000, 700)
    for cuda in (True, False):  # HBM copy / zero-copy alias (UVA), as for `indices`
        q = _quiver(indptr, indices, edge_ids=user, cuda=cuda)
        assert q.has_edge_ids
        out, cnt, eid = q.sample_neighbor(0, _dev(seeds), 7, return_eid=True)
        _, _, r_pos = oracle.sample_neighbor_pos(indptr, indices, seeds, 7)
        assert torch.equal(eid.cpu(), torch.from_numpy(user[r_pos]))


def test_khop_e_id_matches_oracle_and_default_stays_empty(oracle):
    import quiver
    indptr, indices = powerlaw_csr(30000, 25.0, seed=23)
    topo = quiver.CSRTopo(indptr=torch.from_numpy(indptr), indices=torch.from_numpy(indices))
    seeds = np.random.default_rng(2).permutation(30000)[:1024]
    plain = quiver.pyg.GraphSageSampler(topo, [15, 10, 5], device=0, mode="GPU")
    n_id0, _, adjs0 = plain.sample(torch.from_numpy(seeds))
    assert all(a.e_id.numel() == 0 and not a.e_id.is_cuda for a in adjs0)  # the reference's contract (sage_sampler.py:143)
    for fused in (True, False):
        s = quiver.pyg.GraphSageSampler(topo, [15, 10, 5], device=0, mode="GPU", return_eid=True)
        s.fused = fused
        n_id, bs, adjs = s.sample(torch.from_numpy(seeds))
        o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, [15, 10, 5], with_eid=True)
        assert torch.equal(n_id, n_id0) and torch.equal(n_id.cpu(), torch.from_numpy(o_nid))
        for adj, adj0, (o_ei, o_size, o_pos) in zip(adjs, adjs0, o_adjs):
            assert torch.equal(adj.edge_index, adj0.edge_index)
            assert torch.equal(adj.edge_index.cpu(), torch.from_numpy(o_ei))
            assert adj.e_id.is_cuda and torch.equal(adj.e_id.cpu(), torch.from_numpy(o_pos)), fused
            # PyG's invariant: edge e_id[e] of the graph runs from n_id[edge_index[0, e]] (its column) to the target
            src_global = n_id[adj.edge_index[0]].cpu().numpy()
            assert np.array_equal(indices[adj.e_id.cpu().numpy()], src_global)


def test_khop_e_id_with_mega_row(oracle):
    """A 150 k-degree row goes through the chain-splitting path (segment workers + global reservoir): its positions too."""
    rng = np.random.default_rng(5)
    n = 200000
    deg = rng.integers(0, 8, n)
    deg[7] = 150000
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(0, n, int(indptr[-1]), dtype=np.int64)
    q = _quiver(indptr, indices)
    others = rng.permutation(n)[:300]
    seeds = np.concatenate([[7], others[others != 7]])
    n_id, hops = q.sample_khop(_dev(seeds), [10, 5], with_eid=True)
    pos0 = hops[0][3].cpu().numpy()
    _, _, r_pos = oracle.sample_neighbor_pos(indptr, indices, seeds, 10)
    assert np.array_equal(pos0, r_pos)


def test_sample_sub_is_one_call_and_equals_the_two_calls(oracle):
    from torch_quiver import _lib
    indptr, indices = powerlaw_csr(20000, 15.0, seed=24)
    q = _quiver(indptr, indices)
    seeds = np.random.default_rng(3).permutation(20000)[:5000]
    for k in (3, 10, 40):
        out, cnt = q.sample_neighbor(0, _dev(seeds), k)
        want = q.reindex_single(_dev(seeds), out, cnt)
        before = _lib.launch_count()
        got = q.sample_sub(0, _dev(seeds), k)
        launches = _lib.launch_count() - before
        for a, b in zip(got, want):
            assert torch.equal(a, b), k
        o_out, o_cnt = oracle.sample_neighbor(indptr, indices, seeds, k)
        o_f, o_row, o_col = oracle.reindex(seeds, o_out, o_cnt)
        assert torch.equal(got[0].cpu(), torch.from_numpy(o_f)) and torch.equal(got[1].cpu(), torch.from_numpy(o_row))
        assert torch.equal(got[2].cpu(), torch.from_numpy(o_col))
        assert launches <= 6  # one fused one-hop qv_khop (count, sample, insert, scan, emit), not two host round trips
    # k = -1 has no static bound: the two-call path serves it
    f, row, col = q.sample_sub(0, _dev(seeds[:200]), -1)
    o_out, o_cnt = oracle.sample_neighbor(indptr, indices, seeds[:200], -1)
    o_f, o_row, o_col = oracle.reindex(seeds[:200], o_out, o_cnt)
    assert torch.equal(f.cpu(), torch.from_numpy(o_f)) and torch.equal(col.cpu(), torch.from_numpy(o_col))


def test_fast_mode_e_id_is_structurally_valid():
    indptr, indices = powerlaw_csr(8000, 40.0, seed=25)
    q = _quiver(indptr, indices)
    q.set_fast(True)
    seeds = np.random.default_rng(4).permutation(8000)[:2000]
    n_id, hops = q.sample_khop(_dev(seeds), [10, 5], with_eid=True)
    frontier = seeds
    for h, (edge_index, n_src, n_dst, eid) in enumerate(hops):
        pos = eid.cpu().numpy()
        tgt = edge_index[1].cpu().numpy()
        nodes = n_id.cpu().numpy()[:n_dst]
        assert np.all(pos >= indptr[nodes[tgt]]) and np.all(pos < indptr[nodes[tgt] + 1])  # inside the target's row
        assert np.array_equal(indices[pos], n_id.cpu().numpy()[edge_index[0].cpu().numpy()])
        key = tgt.astype(np.int64) * (indices.shape[0] + 1) + pos
        assert np.unique(key).shape[0] == key.shape[0]  # no position twice for one target: without replacement
