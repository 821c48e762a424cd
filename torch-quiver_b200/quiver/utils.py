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
    """COO -> CSR exactly as the reference does it (utils.py:109-116): scipy merges duplicate edges, sorts the
    columns of each row and infers rows = max(src) + 1."""
    from scipy.sparse import csr_matrix
    src = edge_index[0].numpy()
    dst = edge_index[1].numpy()
    data = np.zeros(dst.shape, dtype=np.int32)
    return csr_matrix((data, (src, dst)))


class CSRTopo:
    """Graph topology in CSR format (reference: utils.py:119-226).

    >>> csr_topo = CSRTopo(edge_index=edge_index)
    >>> csr_topo = CSRTopo(indptr=indptr, indices=indices)
    """

    def __init__(self, edge_index=None, indptr=None, indices=None, eid=None):
        if edge_index is not None:
            m = get_csr_from_coo(edge_index)
            self.indptr_ = torch.from_numpy(m.indptr).type(torch.long)
            self.indices_ = torch.from_numpy(m.indices).type(torch.long)
        elif indptr is not None and indices is not None:
            if isinstance(indptr, np.ndarray):
                indptr, indices = torch.from_numpy(indptr), torch.from_numpy(indices)
            self.indptr_ = indptr.type(torch.long)
            self.indices_ = indices.type(torch.long)
        else:
            raise ValueError("CSRTopo needs edge_index or (indptr, indices)")
        self.eid_ = eid
        self.feature_order_ = None

    @property
    def indptr(self):
        return self.indptr_

    @property
    def indices(self):
        return self.indices_

    @property
    def eid(self):
        return self.eid_

    @property
    def feature_order(self):
        return self.feature_order_

    @feature_order.setter
    def feature_order(self, feature_order):
        self.feature_order_ = feature_order

    @property
    def degree(self):
        return self.indptr[1:] - self.indptr[:-1]

    @property
    def node_count(self):
        return self.indptr_.shape[0] - 1

    @property
    def edge_count(self):
        return self.indices_.shape[0]

    def share_memory_(self):
        self.indptr_.share_memory_()
        self.indices_.share_memory_()
        if self.eid_ is not None:
            self.eid_.share_memory_()
        if self.feature_order_ is not None:
            self.feature_order_.share_memory_()


def reindex_by_config(adj_csr: CSRTopo, graph_feature, gpu_portion):
    """Degree-descending row order with the hot `gpu_portion` prefix shuffled so clique shards are load balanced
    (reference: utils.py:229-241).  Returns (permuted feature, new_order) with feature_new[new_order[i]] == feature[i]."""
    node_count = adj_csr.indptr.shape[0] - 1
    hot = int(node_count * gpu_portion)
    degree = adj_csr.indptr[1:] - adj_csr.indptr[:-1]
    _, prev_order = torch.sort(degree, descending=True)
    prev_order[:hot] = prev_order[torch.randperm(hot)]
    new_order = torch.zeros_like(prev_order)
    new_order[prev_order] = torch.arange(node_count, dtype=torch.long)
    return graph_feature[prev_order], new_order


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
