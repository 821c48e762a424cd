from .sage_sampler import Adj, GraphSageSampler

__all__ = ["GraphSageSampler", "Adj"]
