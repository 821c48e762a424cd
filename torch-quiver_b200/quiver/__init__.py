"""`quiver` -- host-side mirror of the reference's Python API for the sampler + feature-gather hot path
(reference: srcs/python/quiver/__init__.py:2-11).  Same names, argument meaning and return shapes, so PyG training
loops written against torch-quiver run unchanged; every byte of device work goes through libquiver_b200.so.

In scope: CSRTopo, pyg.GraphSageSampler, Feature, ShardTensor, p2pCliqueTopo, init_p2p, mp.spawn pickling.
Next-tier rows also mirrored: sample_prob / cal_neighbor_prob and the access-probability partitioner (partition.py).
Out of scope: serving, multi-host NcclComm / DistFeature, MixedGraphSageSampler.
"""
from . import multiprocessing  # noqa: F401  (registers the ForkingPickler reducers, as the reference does)
from . import pyg
from .feature import DeviceConfig, Feature
from .partition import load_quiver_feature_partition, quiver_partition_feature
from .pyg import GraphSageSampler
from .shard_tensor import ShardTensor, ShardTensorConfig
from .utils import CSRTopo, init_p2p, parse_size
from .utils import Topo as p2pCliqueTopo

__all__ = ["Feature", "DeviceConfig", "GraphSageSampler", "CSRTopo", "p2pCliqueTopo", "init_p2p", "ShardTensor",
           "ShardTensorConfig", "parse_size", "pyg", "quiver_partition_feature", "load_quiver_feature_partition"]
