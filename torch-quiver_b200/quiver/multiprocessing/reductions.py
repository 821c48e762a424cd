"""ForkingPickler reducers so Feature / GraphSageSampler can be passed through torch.multiprocessing.spawn
(reference: srcs/python/quiver/multiprocessing/reductions.py:5-33)."""
from multiprocessing.reduction import ForkingPickler


def rebuild_feature(ipc_handle):
    from ..feature import Feature
    return Feature.lazy_from_ipc_handle(ipc_handle)


def reduce_feature(feature):
    return rebuild_feature, (feature.share_ipc(), )


def rebuild_pyg_sampler(cls, ipc_handle):
    return cls.lazy_from_ipc_handle(ipc_handle)


def reduce_pyg_sampler(sampler):
    return rebuild_pyg_sampler, (type(sampler), sampler.share_ipc())


def init_reductions():
    from ..feature import Feature
    from ..pyg.sage_sampler import GraphSageSampler
    ForkingPickler.register(Feature, reduce_feature)
    ForkingPickler.register(GraphSageSampler, reduce_pyg_sampler)
