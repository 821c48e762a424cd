"""NVLink bytes of the gather kernel FROM HARDWARE COUNTERS (north_star: "per-kernel ncu captures of ... NVLink GB/s").
One process, two GPUs (ncu cannot follow a multi-rank job): GPU 0 gathers random 1 KiB rows from a table that is half
local HBM, half GPU 1's HBM (a plain peer pointer -- measured equal to a CUDA-IPC mapping, profiles/r2_peer_probe_table_size.txt).
    ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum -k regex:gather -c 3 python profiles/nvlink_counters.py
The script also prints the event-timed rate so the counter figure can be set against it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200")):
    sys.path.insert(0, p)
import torch
import torch_quiver as qv

D, n_idx, GB = 256, 400_000, 8
rows = GB * (1 << 30) // (D * 4)
torch.cuda.set_device(0)
qv.init_p2p([0, 1])
st = qv.ShardTensor(0)
st.append_empty(rows, [D], torch.float32, 0)
st.append_empty(rows, [D], torch.float32, 1)
g = torch.Generator(device="cuda").manual_seed(1)
idx = [torch.randint(0, 2 * rows, (n_idx, ), generator=g, device="cuda") for _ in range(4)]
out = torch.empty(n_idx, D, device="cuda")
st.gather(idx[0], out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in idx[1:]:
    st.gather(i, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
peer_rows = sum(int((i >= rows).sum()) for i in idx[1:]) / 3
print(f"rows per launch {n_idx}, of them on the peer {peer_rows:.0f}; {ms:.3f} ms per launch; peer bytes {peer_rows * D * 4 / 1e6:.1f} MB "
      f"=> {peer_rows * D * 4 / ms / 1e6:.1f} GB/s over NVLink (event-timed)")
