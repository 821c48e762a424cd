"""CSR container, P2P clique topology, size parsing, degree-ordered feature placement.
Reference: srcs/python/quiver/utils.py."""
from typing import List

import numpy as np
import torch

import torch_quiver as torch_qv


def _maximal_cliques(adj, nodes):
    """Bron-Kerbosch without pivoting; order of discovery follows node order (reference: utils.py:7-32)."""
    found = []

    def grow(clique, cand, excl):
        if not cand and not excl:
            found.append(clique)
            return
        for v in list(cand):
            grow(clique + [v], [u for u in cand if adj[v][u]], [u for u in excl if adj[v][u]])
            cand.remove(v)
            excl.append(v)

    grow([], list(nodes), [])
    return found


def color_mat(access_book, device_list):
    """Assign every device to the first maximal P2P clique that contains it (reference: utils.py:35-50).

    The reference hard-codes 8 devices to [[0,1,2,3],[4,5,6,7]] (utils.py:40-41, a DGX-1 assumption); on NVSwitch
    machines every pair is peer-accessible, so detection yields ONE clique of 8 -- which is what B200 nodes are."""
    device2clique = dict.fromkeys(device_list, -1)
    clique2device = {}
    cliques = _maximal_cliques(access_book, range(len(device_list)))
    cid = 0
    for clique in cliques:
        members = [device_list[i] for i in clique if device2clique[device_list[i]] == -1]
        if not members:
            continue
        clique2device[cid] = members
        for d in members:
            device2clique[d] = cid
        cid += 1
    return device2clique, clique2device


class Topo:
    """P2P access topology of a device list (reference: utils.py:53-106)."""

    def __init__(self, device_list: List[int]) -> None:
        n = len(device_list)
        access = [[0] * n for _ in range(n)]
        for i, src in enumerate(device_list):
            for j, dst in enumerate(device_list):
                if i != j and torch_qv.can_device_access_peer(src, dst):
                    access[i][j] = access[j][i] = 1
        self.Device2p2pClique, self.p2pClique2Device = color_mat(access, list(device_list))

    def get_clique_id(self, device_id: int):
        return self.Device2p2pClique[device_id]

    def info(self):
        return "".join(f"Devices {devs} support p2p access with each other\n"
                       for devs in self.p2pClique2Device.values())

    @property
    def p2p_clique(self):
        return self.p2pClique2Device


def get_csr_from_coo(edge_index):
    """COO -> CSR with the reference's semantics (utils.py:109-116: `scipy.sparse.csr_matrix((zeros, (row, col)))`):
    duplicate edges are merged, the columns of each row come out sorted, and the row count is max(src) + 1 -- so a node
    that only ever appears as a destination has no row."""
    from scipy.sparse import coo_matrix
    rows, cols = (edge_index[i].numpy() for i in (0, 1))
    return coo_matrix((np.zeros(cols.shape[0], dtype=np.int32), (rows, cols))).tocsr()


def _as_long(t):
    return (torch.from_numpy(t) if isinstance(t, np.ndarray) else t).to(torch.long)


class CSRTopo:
    """Graph topology in CSR format (reference: utils.py:119-226).

    >>> csr_topo = CSRTopo(edge_index=edge_index)
    >>> csr_topo = CSRTopo(indptr=indptr, indices=indices)

    `indptr`, `indices`, `eid` are read-only views of what was passed in (CPU int64 tensors); `feature_order` is filled in
    by Feature.from_cpu_tensor when rows get re-ordered by degree.
    """
    _SHARED = ("indptr_", "indices_", "eid_", "feature_order_")

    def __init__(self, edge_index=None, indptr=None, indices=None, eid=None):
        if edge_index is not None:
            csr = get_csr_from_coo(edge_index)
            indptr, indices = csr.indptr, csr.indices
        elif indptr is None or indices is None:
            raise ValueError("CSRTopo needs edge_index or (indptr, indices)")
        self.indptr_, self.indices_ = _as_long(indptr), _as_long(indices)
        self.eid_ = eid
        self.feature_order_ = None

    indptr = property(lambda self: self.indptr_)
    indices = property(lambda self: self.indices_)
    eid = property(lambda self: self.eid_)
    degree = property(lambda self: torch.diff(self.indptr_))
    node_count = property(lambda self: self.indptr_.numel() - 1)
    edge_count = property(lambda self: self.indices_.numel())

    @property
    def feature_order(self):
        return self.feature_order_

    @feature_order.setter
    def feature_order(self, order):
        self.feature_order_ = order

    def share_memory_(self):
        for name in self._SHARED:
            t = getattr(self, name)
            if t is not None:
                t.share_memory_()


def reindex_by_config(adj_csr: CSRTopo, graph_feature, gpu_portion):
    """Degree-descending row order with the hot `gpu_portion` prefix shuffled so clique shards are load balanced
    (reference: utils.py:229-241).  Returns (permuted feature, new_order) with feature_new[new_order[i]] == feature[i]."""
    n = adj_csr.node_count
    by_degree = torch.argsort(adj_csr.degree.cpu(), descending=True, stable=False)  # (a CSRTopo may hold device tensors)
    hot = int(n * gpu_portion)
    by_degree[:hot] = by_degree[:hot][torch.randperm(hot)]
    inverse = torch.empty_like(by_degree)
    inverse[by_degree] = torch.arange(n, dtype=torch.long)
    return graph_feature[by_degree], inverse


def reindex_feature(graph: CSRTopo, feature, ratio):
    assert isinstance(graph, CSRTopo), "Input graph should be CSRTopo object"
    return reindex_by_config(graph, feature, ratio)


def init_p2p(device_list: List[int]):
    """Enable peer access between the devices of `device_list` (reference: utils.py:250-256)."""
    torch_qv.init_p2p(device_list)


UNITS = {"KB": 2**10, "MB": 2**20, "GB": 2**30, "K": 2**10, "M": 2**20, "G": 2**30}


def parse_size(sz) -> int:
    """"0.9M" / "3GB" / int / float -> bytes (reference: utils.py:259-280)."""
    if isinstance(sz, int):
        return sz
    if isinstance(sz, float):
        return int(sz)
    if isinstance(sz, str):
        up = sz.upper()
        for suf in sorted(UNITS, key=len, reverse=True):
            if up.endswith(suf):
                return int(float(sz[:-len(suf)]) * UNITS[suf])
    raise Exception("invalid size: {}".format(sz))
