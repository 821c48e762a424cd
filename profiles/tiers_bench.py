"""Gather bandwidth per tier: local HBM / peer HBM over NVLink / pinned host over PCIe, both kernel variants.
Single process driving device 0 (needs >= 2 GPUs for the peer rows).  Tuning aid; bench.py is the judged number."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch_quiver as qv
from microbench import time_ms

res = []
torch.cuda.set_device(0)
n_gpu = torch.cuda.device_count()
if n_gpu > 1:
    qv.init_p2p([0, 1])
for d in (100, 256, 768):
    rows, n = 1_000_000, 600_000
    x = torch.rand(rows, d)
    rb = d * 4
    layouts = {"local": [(x, 0)], "host": [(x, -1)]}
    if n_gpu > 1:
        layouts["peer"] = [(x, 1)]
        layouts["half_peer"] = [(x[:rows // 2], 0), (x[rows // 2:], 1)]
    for name, parts in layouts.items():
        st = qv.ShardTensor(0)
        for t, dev in parts:
            st.append(t.clone() if dev == -1 else t, dev)
        m = n if name != "host" else 100_000
        idxs = [torch.randint(0, rows, (m, ), device="cuda") for _ in range(4)]
        for variant in (1, 2):
            st.gather_variant = variant
            i = [0]
            def run():
                st[idxs[i[0] % 4]]; i[0] += 1
            ms = time_ms(run, reps=6, warm=2)
            frac_remote = {"local": 0, "host": 1, "peer": 1, "half_peer": 0.5}[name]
            r = dict(d=d, row_bytes=rb, tier=name, variant=variant, ms=round(ms, 4), out_GBps=round(m * rb / ms / 1e6, 1),
                     remote_GBps=round(m * rb * frac_remote / ms / 1e6, 1))
            res.append(r); print(r, flush=True)
        del st
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tiers.json"), "w"), indent=1)
