"""CPU suite, part 4: the one-process-per-GPU placement logic (quiver.shard_tensor.build_from_ranks) under a real
world_size-2 `gloo` rendezvous -- handle exchange order, row offsets, host tier -- with the device layer faked."""
import os
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


class _FakeItem:
    def __init__(self):
        self.ipc = None

    def from_ipc(self, ipc):
        self.ipc = ipc
        self.shape = ipc[3]
        self.device = ipc[0]

    def share_ipc(self):
        return self.ipc


class _FakeST:
    """Records what would be placed where; `share_ipc` fabricates a handle that encodes (device, rows)."""

    def __init__(self, device):
        self.device_, self.parts = device, []

    def append(self, t, dev=None):
        if isinstance(t, _FakeItem):
            self.parts.append(("ipc", t.ipc[0], t.ipc[3][0], bytes(t.ipc[2])))
        else:
            self.parts.append(("local" if dev is not None and dev >= 0 else "host", dev, t.shape[0], t))

    def adopt(self, other):
        self.parts.append(("adopted",) + other.parts[0][1:])
        other.parts.clear()

    def size(self, dim):
        return sum(p[2] for p in self.parts)

    def share_ipc(self):
        kind, dev, rows, t = self.parts[0]
        item = _FakeItem()
        item.ipc = (dev, 4, f"handle-dev{dev}-rows{rows}".encode().ljust(64, b"\0"), [rows, t.shape[1]])
        return [item]


def _worker(rank, world, port, results):
    sys.path[:0] = [ROOT, PKG]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import quiver.shard_tensor as qst
    fake = types.SimpleNamespace(ShardTensor=_FakeST, ShardTensorItem=_FakeItem, can_device_access_peer=lambda a, b: True,
                                 init_p2p=lambda d: None)
    qst.torch_qv = fake
    rows = 100 + 50 * rank  # ragged shards
    local = torch.full((rows, 8), float(rank))
    cold = torch.zeros(7, 8) if rank == 0 else None
    st = qst.build_from_ranks(local, device=rank, cpu_part=cold)
    parts = st.shard_tensor.parts
    ok = [p[0] for p in parts[:world]] == ["adopted" if r == rank else "ipc" for r in range(world)]
    ok &= [p[2] for p in parts[:world]] == [100 + 50 * r for r in range(world)]
    ok &= all(p[0] != "ipc" or p[3].startswith(f"handle-dev{p[1]}-rows{p[2]}".encode()) for p in parts)
    offs = st.shard_tensor_config.tensor_offset_device
    ok &= [(offs[r].start, offs[r].end) for r in range(world)] == [(0, 100), (100, 250)][:world]
    ok &= (len(parts) == world + 1 and parts[-1][0] == "host") if rank == 0 else len(parts) == world
    results[rank] = 1 if ok else 0
    dist.destroy_process_group()


def test_build_from_ranks_world2_gloo():
    world = 2
    results = torch.zeros(world, dtype=torch.int32).share_memory_()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert results.tolist() == [1, 1]


def _tier_worker(rank, world, port, results):
    sys.path[:0] = [ROOT, PKG]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quiver.shard_tensor import tier_ranges
    n, hot, cold = 1003, 200, 101
    mine = tier_ranges(n, hot, cold, world, rank)
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    ok = all(e["hot"] == (0, hot) and e["cold"] == (n - cold, n) and e["striped"] == (hot, n - cold) for e in everyone)
    stripes = [e["stripe"] for e in everyone]
    ok &= stripes[0][0] == hot and stripes[-1][1] == n - cold
    ok &= all(stripes[r][1] == stripes[r + 1][0] for r in range(world - 1))  # the ranks' blocks tile [hot, n - cold) in order
    ok &= mine == everyone[rank]
    results[rank] = 1 if ok else 0
    dist.destroy_process_group()


def test_tiered_layout_ranges_world2_gloo():
    """build_tiered_inplace's layout arithmetic (hot prefix replicated | one stripe per rank | cold host suffix) agreed
    on by two ranks over a real gloo rendezvous; plus the single-rank and error cases."""
    from quiver.shard_tensor import tier_ranges
    world = 2
    results = torch.zeros(world, dtype=torch.int32).share_memory_()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_tier_worker, args=(world, port, results), nprocs=world, join=True)
    assert results.tolist() == [1, 1]
    one = tier_ranges(1000, 300, 100, 1, 0)  # a single GPU has no replicated tier: everything that is not cold is its shard
    assert one["hot"] == (0, 0) and one["stripe"] == (0, 900) and one["cold"] == (900, 1000)
    eight = [tier_ranges(100_000_000, 40_000_000, 0, 8, r)["stripe"] for r in range(8)]
    assert eight[0] == (40_000_000, 47_500_000) and eight[-1][1] == 100_000_000
    with pytest.raises(ValueError):
        tier_ranges(10, 8, 5, 2, 0)
