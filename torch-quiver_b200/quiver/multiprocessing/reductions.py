"""ForkingPickler reducers so Feature / GraphSageSampler can be passed through torch.multiprocessing.spawn
(reference: srcs/python/quiver/multiprocessing/reductions.py:5-33).

Both classes travel the same way: `share_ipc()` in the parent produces a picklable handle, `lazy_from_ipc_handle(handle)`
in the child rebuilds a lazily-initialised object bound to the child's current CUDA device.  One reducer serves every
registered class."""
from multiprocessing.reduction import ForkingPickler


def _revive(cls, handle):
    return cls.lazy_from_ipc_handle(handle)


def _reduce(obj):
    return _revive, (type(obj), obj.share_ipc())


def init_reductions():
    from ..feature import Feature
    from ..pyg.sage_sampler import GraphSageSampler
    for cls in (Feature, GraphSageSampler):
        ForkingPickler.register(cls, _reduce)
