"""Peer-tier gather rate vs TABLE SIZE and mapping kind (2 GPUs).  Why: the north-star table striped over N GPUs puts tens of
GB behind each peer mapping; round 1 measured the peer tier on a < 1 GB table only.
  torchrun --nproc-per-node 2 profiles/peer_probe.py     (rank 0 gathers; rank 1 owns the IPC-exported table)
Prints GB/s (output bytes) for random 1 KiB rows out of a peer table of 1 / 8 / 40 GB:
  same-process peer pointer (cudaDeviceEnablePeerAccess)   vs   cross-process CUDA IPC mapping (cudaIpcOpenMemHandle)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import torch_quiver as qv

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
D, n_idx = 256, 400_000
row_bytes = D * 4


def bench(st, rows, label):
    g = torch.Generator(device="cuda").manual_seed(1)
    idx = [torch.randint(0, rows, (n_idx, ), generator=g, device="cuda") for _ in range(6)]
    out = torch.empty(n_idx, D, device="cuda")
    st.gather(idx[0], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in idx[1:]:
        st.gather(i, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label:60s} {rows * row_bytes / 1e9:6.1f} GB table: {n_idx * row_bytes / ms / 1e6:8.1f} GB/s  ({ms:.3f} ms)", flush=True)


qv.init_p2p([0, 1])
for gb in (1, 8, 40):
    rows = gb * (1 << 30) // row_bytes
    # ---- cross-process: rank 1 allocates and exports, rank 0 maps through CUDA IPC -----------------------------------
    local = qv.ShardTensor(rank)
    handle = None
    if rank == 1:
        v = local.append_empty(rows, [D], torch.float32, 1)
        v[:: max(1, rows // 4096)].fill_(1.0)
        handle = local.share_ipc()[0].share_ipc()
    box = [handle]
    dist.broadcast_object_list(box, src=1)
    if rank == 0:
        st = qv.ShardTensor(0)
        item = qv.ShardTensorItem()
        item.from_ipc(box[0])
        st.append(item)
        bench(st, rows, "cross-process CUDA IPC mapping of GPU1 memory")
        del st
    dist.barrier()
    del local
    torch.cuda.synchronize()
    dist.barrier()
    # ---- same process: rank 0 allocates on GPU 1 and reads it through a plain peer pointer ---------------------------
    if rank == 0:
        st = qv.ShardTensor(0)
        st.append_empty(rows, [D], torch.float32, 1)
        bench(st, rows, "same-process peer pointer to GPU1 memory")
        del st
        st = qv.ShardTensor(0)
        st.append_empty(rows, [D], torch.float32, 0)
        bench(st, rows, "local HBM")
        del st
    dist.barrier()
dist.destroy_process_group()
