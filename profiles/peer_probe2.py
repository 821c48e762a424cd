"""Why does the N=2 north-star gather collapse to ~100 GB/s on peer rows when round 1's small table reached 500+?
Separates the candidates on 2 GPUs (random 1 KiB rows, 400 k per launch, 20 GB per GPU):
  A  rank 0 reads rank 1's memory, rank 1 idle                 (baseline: 737 GB/s)
  B  both ranks read each other's memory at the same time      (bidirectional link load)
  C  rank 0 reads rank 1's memory while rank 1 gathers from its OWN memory (remote reads into a busy HBM)
  D  rank 0 reads a table that is half local, half peer, rank 1 idle (mixed tiers inside one kernel)
  E  D on both ranks at the same time
  F  like E, but the peer-read region is the one its owner gathers from (what bench.py does at N=2)
  torchrun --nproc-per-node 2 profiles/peer_probe2.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import torch_quiver as qv

rank = int(os.environ.get("RANK", 0))
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
D, n_idx, GB = 256, 400_000, int(os.environ.get("QV_PROBE_GB", "20"))
row_bytes = D * 4
rows = int(os.environ.get("QV_PROBE_ROWS", GB * (1 << 30) // row_bytes))  # e.g. 50000000: NOT a multiple of 2 MiB
qv.init_p2p([0, 1])
# QV_MALLOC_GRANULE (read by qv_malloc) = allocation granularity under test; 1 = exact sizes (the driver's default)
local = qv.ShardTensor(rank)
local.append_empty(rows, [D], torch.float32, rank)
box = [None, None]
dist.all_gather_object(box, local.share_ipc()[0].share_ipc())
peer = qv.ShardTensor(rank)
item = qv.ShardTensorItem()
item.from_ipc(box[1 - rank])
peer.append(item)
mixed = qv.ShardTensor(rank)  # [my rows | the peer's rows]
mixed.adopt(local) if False else None
loc2 = qv.ShardTensor(rank)
loc2.append_empty(rows, [D], torch.float32, rank)
mixed.adopt(loc2)
item2 = qv.ShardTensorItem()
item2.from_ipc(box[1 - rank])
mixed.append(item2)
g = torch.Generator(device="cuda").manual_seed(1 + rank)
idx1 = [torch.randint(0, rows, (n_idx, ), generator=g, device="cuda") for _ in range(6)]
idx2 = [torch.randint(0, 2 * rows, (n_idx, ), generator=g, device="cuda") for _ in range(6)]
out = torch.empty(n_idx, D, device="cuda")


def timed(st, idx, active):
    dist.barrier()
    torch.cuda.synchronize()
    if not active:
        dist.barrier()
        return None
    st.gather(idx[0], out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        for i in idx[1:]:
            st.gather(i, out=out)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    return e0.elapsed_time(e1) / 20


def report(label, ms, remote_frac):
    if ms is not None:
        print(f"rank {rank} {label:58s} {ms:.3f} ms  out {n_idx * row_bytes / ms / 1e6:7.1f} GB/s  peer-rows {remote_frac * n_idx * row_bytes / ms / 1e6:7.1f} GB/s",
              flush=True)


report("A  remote gather, other rank idle", timed(peer, idx1, rank == 0), 1.0)
report("B  remote gather, both ranks at once", timed(peer, idx1, True), 1.0)
report("C  rank0 remote / rank1 local at once", timed(peer if rank == 0 else local, idx1, True), 1.0 if rank == 0 else 0.0)
report("D  half local half peer, other rank idle", timed(mixed, idx2, rank == 0), 0.5)
report("E  half local half peer, both ranks at once", timed(mixed, idx2, True), 0.5)
# F: like bench.py -- the region a rank reads over NVLink is the SAME region its owner is gathering from locally
mixed2 = qv.ShardTensor(rank)
mixed2.adopt(local)
item3 = qv.ShardTensorItem()
item3.from_ipc(box[1 - rank])
mixed2.append(item3)
report("F  [own shard | peer's own shard], both ranks at once", timed(mixed2, idx2, True), 0.5)
report("F' same, other rank idle", timed(mixed2, idx2, rank == 0), 0.5)
dist.destroy_process_group()
