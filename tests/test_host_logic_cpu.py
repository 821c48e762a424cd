"""CPU suite, part 3: host-side logic of the `quiver` mirror (placement arithmetic, CSR ingestion, topology,
pickling) checked against the reference's documented behaviour, with the device layer replaced by a recording fake."""
import pickle
import types

import numpy as np
import pytest
import torch

import quiver
from quiver import utils as qutils
from graphs import powerlaw_csr


class FakeShardTensor:
    """Stands in for torch_quiver.ShardTensor: records placement, answers gathers with plain indexing (TEST ONLY)."""

    def __init__(self, device):
        self.device_, self.parts, self.devices = device, [], []

    def append(self, t, dev=None):
        self.parts.append(t)
        self.devices.append(dev)

    def size(self, dim):
        return sum(p.shape[0] for p in self.parts) if dim == 0 else self.parts[0].shape[dim]

    def shape(self):
        return [self.size(0)] + list(self.parts[0].shape[1:])

    def gather(self, idx, feature_order=None):
        full = torch.cat(self.parts)
        idx = idx if feature_order is None else feature_order[idx]
        return full[idx]

    def __getitem__(self, idx):
        return self.gather(idx)


@pytest.fixture
def fake_device_layer(monkeypatch):
    fake = types.SimpleNamespace(ShardTensor=FakeShardTensor, ShardTensorItem=object,
                                 can_device_access_peer=lambda a, b: True, init_p2p=lambda devs: None)
    monkeypatch.setattr("quiver.shard_tensor.torch_qv", fake)
    monkeypatch.setattr("quiver.utils.torch_qv", fake)
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: self, raising=True)
    return fake


def test_parse_size():
    assert quiver.parse_size("0.9M") == int(0.9 * 2**20)
    assert quiver.parse_size("3GB") == 3 * 2**30
    assert quiver.parse_size("200K") == 200 * 1024
    assert quiver.parse_size(12345) == 12345 and quiver.parse_size(1.5e3) == 1500
    with pytest.raises(Exception):
        quiver.parse_size("lots")


def test_csrtopo_from_coo_matches_scipy_semantics():
    # duplicate edges merged, columns sorted, rows = max(src)+1
    ei = torch.tensor([[2, 0, 0, 2, 0], [1, 3, 1, 1, 3]])
    topo = quiver.CSRTopo(edge_index=ei)
    assert topo.indptr.tolist() == [0, 2, 2, 3]
    assert topo.indices.tolist() == [1, 3, 1]
    assert topo.node_count == 3 and topo.edge_count == 3
    assert topo.degree.tolist() == [2, 0, 1]
    indptr, indices = powerlaw_csr(50, 4.0, seed=1)
    t2 = quiver.CSRTopo(indptr=indptr, indices=indices)  # numpy input
    assert t2.indptr.dtype == torch.long and t2.node_count == 50


def test_topo_cliques_without_hardcoded_eight(monkeypatch):
    # NVSwitch: every pair is peer-accessible -> ONE clique of 8 (the reference forces [[0-3],[4-7]], utils.py:40-41)
    monkeypatch.setattr("quiver.utils.torch_qv", types.SimpleNamespace(can_device_access_peer=lambda a, b: True))
    t = qutils.Topo(list(range(8)))
    assert t.p2pClique2Device == {0: list(range(8))} and t.get_clique_id(5) == 0
    # two PCIe islands
    monkeypatch.setattr("quiver.utils.torch_qv",
                        types.SimpleNamespace(can_device_access_peer=lambda a, b: (a < 2) == (b < 2)))
    t = qutils.Topo([0, 1, 2, 3])
    assert t.p2pClique2Device == {0: [0, 1], 1: [2, 3]} and "support p2p" in t.info()


def test_reindex_feature_identity():
    # original_feature[ids] == new_feature[new_order[ids]]   (tests/python/cuda/test_graph_reindex.py:58-59)
    indptr, indices = powerlaw_csr(400, 8.0, seed=2)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = torch.arange(400 * 3, dtype=torch.float32).view(400, 3)
    new_x, new_order = qutils.reindex_feature(topo, x, 0.25)
    ids = torch.randint(0, 400, (100, ))
    assert torch.equal(x[ids], new_x[new_order[ids]])
    deg = topo.degree
    assert deg[new_order.argsort()[100:]].tolist() == sorted(deg[new_order.argsort()[100:]].tolist(), reverse=True)


@pytest.mark.parametrize("policy", ["device_replicate", "p2p_clique_replicate"])
def test_feature_placement_arithmetic(fake_device_layer, policy):
    N, D = 1000, 16
    x = torch.arange(N * D, dtype=torch.float32).view(N, D)
    budget = 100 * D * 4  # 100 rows per device
    f = quiver.Feature(rank=0, device_list=[0, 1, 2, 3], device_cache_size=budget, cache_policy=policy)
    assert f.from_cpu_tensor(x) is f
    if policy == "device_replicate":
        # every GPU caches the same 100 hot rows; the launching rank also sees the cold rows (feature.py:219-223,268-273)
        assert sorted(f.device_tensor_list) == [0, 1, 2, 3]
        st = f.device_tensor_list[0].shard_tensor
        assert [p.shape[0] for p in st.parts] == [100, 900] and st.devices == [0, -1]
        assert [p.shape[0] for p in f.device_tensor_list[2].shard_tensor.parts] == [100]
    else:
        # budget x clique size, equal blocks, last GPU takes the remainder, then the host tier (feature.py:204,229-246)
        st = f.clique_tensor_list[0].shard_tensor
        assert [p.shape[0] for p in st.parts] == [100, 100, 100, 100, 600] and st.devices == [0, 1, 2, 3, -1]
    assert f.shape == [N, D] and f.size(0) == N and f.size(1) == D and f.dim() == 2
    idx = torch.randint(0, N, (64, ))
    assert torch.equal(f[idx], x[idx])


def test_feature_zero_cache_and_degree_order(fake_device_layer):
    indptr, indices = powerlaw_csr(300, 6.0, seed=5)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    x = torch.randn(300, 8)
    f = quiver.Feature(rank=0, device_list=[0], device_cache_size=0, cache_policy="device_replicate", csr_topo=topo)
    f.from_cpu_tensor(x)  # device_cache_size = 0 => everything in the host tier (test_features.py:394-398)
    st = f.device_tensor_list[0].shard_tensor
    assert st.devices == [-1] and st.parts[0].shape[0] == 300
    idx = torch.randint(0, 300, (50, ))
    assert torch.equal(f[idx], x[idx])  # the degree permutation is hidden behind feature_order
    assert topo.feature_order is not None


def test_sampler_pickles_lazily():
    indptr, indices = powerlaw_csr(60, 3.0, seed=6)
    topo = quiver.CSRTopo(indptr=indptr, indices=indices)
    from quiver.pyg.sage_sampler import GraphSageSampler, _FakeDevice
    s = GraphSageSampler(topo, [5, 3], _FakeDevice, "GPU")  # what a spawned worker receives
    assert s.quiver is None
    handle = s.share_ipc()
    assert handle[1] == [5, 3] and handle[2] == "GPU"
    s2 = GraphSageSampler.lazy_from_ipc_handle(handle)
    assert s2.quiver is None and s2.sizes == [5, 3]
    from multiprocessing.reduction import ForkingPickler
    blob = ForkingPickler.dumps(s)
    s3 = pickle.loads(blob)
    assert isinstance(s3, GraphSageSampler) and s3.csr_topo.node_count == 60
    with pytest.raises(NotImplementedError):
        GraphSageSampler(topo, [5], "cpu", "CPU")


def _naive_partition(probs, chunk_size):
    """Plain-Python restatement of partition.py:99-161 for checking (distinct scores, no ties)."""
    P, n = len(probs), len(probs[0])
    res = [[] for _ in range(P)]
    start, rot = 0, 0
    while start < n:
        end = min(n, start + chunk_size * P)
        free = set(range(start, end))
        for r_ in range(rot, rot + P):
            r = r_ % P
            score = {v: P * probs[r][v] - sum(probs[q][v] for q in range(P) if q != r) for v in free}
            take = sorted(free, key=lambda v: -score[v])[:chunk_size]
            res[r] += take
            free -= set(take)
        rot += 1
        start = end
    return res


def test_partition_by_access_probability(tmp_path):
    from quiver.partition import (load_quiver_feature_partition, partition_feature_without_replication,
                                  partition_without_replication, quiver_partition_feature, select_nodes)
    g = torch.Generator().manual_seed(0)
    P, n = 3, 1000
    probs = [torch.rand(n, generator=g) for _ in range(P)]
    parts, moved = partition_feature_without_replication(probs, 64, device="cpu")
    allv = torch.cat(parts)
    assert allv.numel() == n and torch.unique(allv).numel() == n  # disjoint cover
    assert max(p.numel() for p in parts) - min(p.numel() for p in parts) <= 64
    naive = _naive_partition([p.tolist() for p in probs], 64)
    assert [sorted(p.tolist()) for p in parts] == [sorted(x) for x in naive]
    # the reference's folder layout round-trips
    book, res, cache = quiver_partition_feature(probs, str(tmp_path / "part"), cache_memory_budget="4K",
                                                per_feature_size=16, chunk_size=64, device="cpu")
    assert book.shape == (n, ) and all(bool((book[res[i]] == i).all()) for i in range(P))
    assert all(c.numel() == (4096 // 16) // P for c in cache)
    assert torch.equal(cache[1], torch.sort(probs[1], descending=True)[1][:85])
    b2, r2, c2 = load_quiver_feature_partition(2, str(tmp_path / "part"))
    assert torch.equal(b2, book) and torch.equal(r2, res[2]) and torch.equal(c2, cache[2])
    with pytest.raises(FileExistsError):
        quiver_partition_feature(probs, str(tmp_path / "part"), device="cpu")
    # id-restricted variant + select_nodes
    ids = torch.randperm(n, generator=g)[:300]
    sub = partition_without_replication("cpu", probs, ids)
    assert sorted(torch.cat(sub).tolist()) == sorted(ids.tolist())
    s, nz = select_nodes("cpu", probs, None)
    assert torch.allclose(s, sum(probs)) and nz.numel() == n
