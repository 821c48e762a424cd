"""Sweep (rows in flight per lane, min blocks per SM) of the flat batched gather on the bench row size.  Tuning aid."""
import os, subprocess, sys
if len(sys.argv) > 1:
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch_quiver as qv
    from microbench import time_ms
    for d in (100, 150):
        rows, n = 2_449_029, 820_000
        x = torch.rand(rows, d)
        st = qv.ShardTensor(0); st.append(x, 0)
        order = torch.randperm(rows, device="cuda")
        idxs = [torch.randint(0, rows, (n,), device="cuda") for _ in range(4)]
        outs = [torch.empty(n, d, device="cuda") for _ in range(2)]
        i = [0]
        def run():
            st.gather(idxs[i[0] % 4], order, out=outs[i[0] % 2]); i[0] += 1
        ms = time_ms(run, reps=20, warm=5)
        print(f"tune={os.environ.get('QV_GATHER_TUNE','0')} d={d}: {ms*1e3:.1f} us  alg {n*(8*d+16)/ms/1e6:.0f} GB/s frac {n*(8*d+16)/ms/1e6/6574.8:.3f}", flush=True)
else:
    for t in ("0", "24", "26", "28", "44", "46", "48", "84", "83", "0"):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, QV_GATHER_TUNE=t))
