"""A/B of the sampling kernel variants (QV_SAMPLE_IMPL bit0 = generic kernel, bit1 = plain `%`).  Tuning aid."""
import os, subprocess, sys
if len(sys.argv) > 1:
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch_quiver as qv
    from bench import SIZES, make_graph
    from microbench import time_ms
    indptr, indices = make_graph(torch.device("cuda"))
    q = qv.device_quiver_from_csr_array(indptr.cpu(), indices, None, 0, True)
    for S in (1024, 16384):
        seeds = [torch.randperm(indptr.numel() - 1, device="cuda")[:S] for _ in range(4)]
        # isolate the sampling kernel: time sample_neighbor's two halves on the biggest hop's frontier
        n_id, hops = q.sample_khop(seeds[0], SIZES[:2])
        outs = []
        def run():
            q.sample_neighbor(0, n_id, SIZES[2])
        ms = time_ms(run, reps=10)
        i = [0]
        def run2():
            q.sample_khop(seeds[i[0] % 4], SIZES); i[0] += 1
        ms2 = time_ms(run2, reps=10)
        print(f"impl={os.environ.get('QV_SAMPLE_IMPL','0')} S={S}: hop3 sample_neighbor({n_id.numel()} seeds) {ms:.3f} ms; khop {ms2:.3f} ms", flush=True)
else:
    for impl in ("0", "1", "2", "3"):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, QV_SAMPLE_IMPL=impl))
