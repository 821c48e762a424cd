"""The drop-in boundary as a BUILT artefact: the reference's pybind11 surface (module.cpp:16-26, quiver_sample.cu:500-513,
quiver_feature.cu:431-473) compiled from torch-quiver_b200/csrc/pybind/torch_quiver_pybind.cpp over the C ABI.  The same
parity checks as the ctypes adapter's suite, through THIS module: sample_neighbor / reindex_single / sample_sub bit-exact
against the oracle, cal_neighbor_prob, the k-hop loop of sage_sampler.py:118-147 written against the reference's binding
names, ShardTensor tiers, CUDA-IPC items, and agreement with the ctypes adapter."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import PKG
from graphs import MINI, powerlaw_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    d = os.path.join(PKG, "torch_quiver_pybind")
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import torch_quiver_pb
    except ImportError as e:
        pytest.fail(f"the pybind adapter is not built (python torch-quiver_b200/csrc/pybind/build.py): {e}")
    return torch_quiver_pb


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).cuda()


def _quiver(pb, indptr, indices, cuda=True):
    return pb.device_quiver_from_csr_array(torch.from_numpy(indptr), torch.from_numpy(indices),
                                           torch.zeros(1, dtype=torch.long), 0, cuda)


@pytest.mark.parametrize("k", [1, 5, 25, 33, 2000])
@pytest.mark.parametrize("S", [1, 64, 65, 1000])
def test_sample_neighbor_and_reindex_bit_exact(pb, oracle, k, S):
    indptr, indices = powerlaw_csr(2000, 30.0, seed=7)
    q = _quiver(pb, indptr, indices)
    seeds = np.random.default_rng(S * 131 + k).permutation(2000)[:S]
    out, cnt = q.sample_neighbor(0, _dev(seeds), k)
    r_out, r_cnt = oracle.sample_neighbor(indptr, indices, seeds, k)
    assert torch.equal(out.cpu(), torch.from_numpy(r_out)) and torch.equal(cnt.cpu(), torch.from_numpy(r_cnt))
    f, row, col = q.reindex_single(_dev(seeds), out, cnt)
    o_f, o_row, o_col = oracle.reindex(seeds, r_out, r_cnt)
    assert torch.equal(f.cpu(), torch.from_numpy(o_f)) and torch.equal(row.cpu(), torch.from_numpy(o_row))
    assert torch.equal(col.cpu(), torch.from_numpy(o_col))
    f2, row2, col2 = q.sample_sub(0, _dev(seeds), k)  # one fused call
    assert torch.equal(f2, f) and torch.equal(row2, row) and torch.equal(col2, col)


def test_mini_fixture_and_uva_mode(pb):
    indptr, indices = np.array(MINI["indptr"]), np.array(MINI["indices"])
    for cuda in (True, False):
        q = _quiver(pb, indptr, indices, cuda=cuda)
        out, cnt = q.sample_neighbor(0, _dev(MINI["seeds"]), MINI["k"])
        assert cnt.tolist() == MINI["counts"] and out[4:6].tolist() == [0, 2]  # seed 1 (the 4th): degree 2 <= k, verbatim
        f, row, col = q.reindex_single(_dev(MINI["seeds"]), _dev(MINI["draw"]), cnt)
        assert f.tolist() == MINI["frontier"] and row.tolist() == MINI["row_idx"] and col.tolist() == MINI["col_idx"]


def test_khop_loop_over_the_reference_binding_names(pb, oracle):
    """GraphSageSampler.sample's loop (sage_sampler.py:118-147) written against sample_neighbor / reindex_single, and the
    one-call extension, both equal to the oracle and to the ctypes adapter."""
    import torch_quiver as qv
    indptr, indices = powerlaw_csr(30000, 25.0, seed=23)
    q = _quiver(pb, indptr, indices)
    q2 = qv.device_quiver_from_csr_array(torch.from_numpy(indptr), torch.from_numpy(indices), None, 0, True)
    seeds = np.random.default_rng(2).permutation(30000)[:1024]
    sizes = [15, 10, 5]
    nodes, adjs = _dev(seeds), []
    for size in sizes:
        out, cnt = q.sample_neighbor(0, nodes, size)
        frontier, row_idx, col_idx = q.reindex_single(nodes, out, cnt)
        adjs.append((torch.stack([col_idx, row_idx]), (frontier.numel(), nodes.numel())))
        nodes = frontier
    o_nid, _, o_adjs = oracle.khop(indptr, indices, seeds, sizes)
    assert torch.equal(nodes.cpu(), torch.from_numpy(o_nid))
    for (ei, size), (o_ei, o_size) in zip(adjs[::-1], o_adjs):
        assert torch.equal(ei.cpu(), torch.from_numpy(o_ei)) and size == tuple(o_size)
    n_id, edge_index, hop_sizes = q.sample_khop(_dev(seeds), sizes)
    n_id2, hops2 = q2.sample_khop(_dev(seeds), sizes)
    assert torch.equal(n_id, nodes) and torch.equal(n_id, n_id2)
    for h in range(3):
        assert torch.equal(edge_index[h], adjs[h][0]) and tuple(hop_sizes[h]) == adjs[h][1]
        assert torch.equal(edge_index[h], hops2[h][0])
    q.rand_seed = 12345
    out, _ = q.sample_neighbor(0, _dev(seeds), 4)
    r_out, _ = oracle.sample_neighbor(indptr, indices, seeds, 4, rand_seed=12345)
    assert torch.equal(out.cpu(), torch.from_numpy(r_out))


def test_cal_neighbor_prob(pb, oracle):
    indptr, indices = powerlaw_csr(5000, 12.0, seed=4)
    q = _quiver(pb, indptr, indices)
    last = np.random.default_rng(0).random(5000).astype(np.float32)
    cur = torch.zeros(5000, device="cuda")
    q.cal_neighbor_prob(0, torch.from_numpy(last).cuda(), cur, 5)
    assert np.array_equal(cur.cpu().numpy(), oracle.cal_next(last, 5, indptr, indices))


@pytest.mark.parametrize("dtype,d", [(torch.float32, 100), (torch.float32, 602), (torch.float16, 256)])
def test_shard_tensor_tiers_and_ipc_items(pb, oracle, dtype, d):
    n = 30000
    x = torch.from_numpy(np.random.default_rng(d).integers(0, 10, (n, d)).astype(np.float32)).to(dtype)
    cold = x[20000:].clone()
    st = pb.ShardTensor(0)
    st.append(x[:12000], 0)
    st.append(x[12000:20000], 0)
    st.append(cold, -1)
    assert st.shape() == [n, d] and st.device_count() == 3 and st.size(0) == n and st.stride(0) == d and st.device() == 0
    idx = torch.from_numpy(np.random.default_rng(1).integers(0, n, 50000)).cuda()
    got = st[idx]
    assert got.dtype == dtype and torch.equal(got.cpu(), x[idx.cpu()])
    if dtype == torch.float32:
        assert np.array_equal(got.cpu().numpy(), oracle.gather([x.numpy()], idx.cpu().numpy()))
    bad = torch.tensor([-1, n, 5], device="cuda")
    rows = st[bad]
    assert bool((rows[:2] == 0).all()) and torch.equal(rows[2].cpu(), x[5])  # invalid ids: zero rows, not stale memory
    items = st.share_ipc()
    assert len(items) == 2
    dev, esz, handle, shape = items[0].share_ipc()
    assert dev == 0 and esz == x.element_size() and len(handle) == 64 and shape == [12000, d]
    clone = pb.ShardTensorItem()
    clone.from_ipc((dev, esz, handle, shape))
    assert clone.share_ipc()[3] == shape
    st.unregister(cold)
    with pytest.raises(RuntimeError):
        st.append(x[:10].cuda(), 0)  # CHECK_CPU, quiver_feature.cu:147


def test_errors_raise_instead_of_exit(pb):
    assert pb.can_device_access_peer(0, 0) and pb.init_p2p([0]) == 0 and pb.abi_version() == 2
    indptr, indices = powerlaw_csr(100, 5.0, seed=1)
    q = _quiver(pb, indptr, indices)
    with pytest.raises(RuntimeError):
        q.sample_neighbor(0, torch.arange(4, dtype=torch.int32, device="cuda"), 3)  # not torch.long
    with pytest.raises(RuntimeError):
        q.sample_neighbor(0, torch.arange(4), 3)  # CPU tensor
