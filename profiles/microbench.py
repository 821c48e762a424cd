"""Kernel-level timings used while tuning (CUDA events, warm, inputs larger than L2).  Not a bench line: bench.py is.
Usage (on the GPU box): python profiles/microbench.py [gather|sampler|all]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import torch_quiver as qv  # noqa: E402

PEAK = 6574.8


def time_ms(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def gather():
    n_rows, n = 2_449_029, 820_000
    out = []
    for d, dtype in [(100, torch.float32), (128, torch.float32), (256, torch.float32), (602, torch.float32),
                     (768, torch.float32), (600, torch.float16)]:
        rows = n_rows if d <= 256 else 600_000
        x = torch.rand(rows, d).to(dtype)
        st = qv.ShardTensor(0)
        st.append(x, 0)
        idxs = [torch.randint(0, rows, (n, ), device="cuda") for _ in range(4)]
        rb = d * x.element_size()
        for variant in (1, 2, 3):
            if variant == 2 and rb % 16:
                continue
            st.gather_variant = variant
            i = [0]

            def run():
                st[idxs[i[0] % 4]]
                i[0] += 1
            ms = time_ms(run)
            gbs = n * (2 * rb + 8) / ms / 1e6
            out.append(dict(d=d, dtype=str(dtype), row_bytes=rb, variant=variant, ms=round(ms, 4), alg_GBps=round(gbs, 1),
                            frac=round(gbs / PEAK, 3), out_GBps=round(n * rb / ms / 1e6, 1)))
            print(out[-1], flush=True)
        del st, x
    return out


def sampler():
    from bench import SIZES, make_graph
    indptr, indices = make_graph(torch.device("cuda"))
    q = qv.device_quiver_from_csr_array(indptr.cpu(), indices, None, 0, True)
    out = []
    for S in (1024, 8192, 65536):
        seeds = [torch.randperm(indptr.numel() - 1, device="cuda")[:S] for _ in range(4)]
        i = [0]

        def run():
            q.sample_khop(seeds[i[0] % 4], SIZES)
            i[0] += 1
        ms = time_ms(run, reps=8)
        n_id, hops = q.sample_khop(seeds[0], SIZES)
        edges = sum(h[0].shape[1] for h in hops)
        out.append(dict(S=S, sizes=SIZES, ms=round(ms, 4), edges=edges, seps=round(edges / ms * 1e3)))
        print(out[-1], flush=True)
    return out


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = {}
    if what in ("gather", "all"):
        res["gather"] = gather()
    if what in ("sampler", "all"):
        res["sampler"] = sampler()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)
