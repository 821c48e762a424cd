"""Does the sampler (private high-priority stream) really run concurrently with a feature gather on the current stream?
Times the k-hop sampler alone, then while 4 back-to-back gathers (~0.5 ms) occupy the current stream.
Usage: python profiles/overlap_probe.py   (one GPU; prints a JSON object)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-quiver_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import bench  # noqa: E402
import quiver  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    indptr, indices = bench.make_graph(dev)
    topo = quiver.CSRTopo(indptr=indptr.cpu(), indices=indices.cpu())
    del indptr, indices
    sampler = quiver.pyg.GraphSageSampler(topo, bench.SIZES, device=0, mode="GPU")
    x = torch.rand(bench.N_NODES, bench.FEAT_DIM)
    feature = quiver.Feature(rank=0, device_list=[0], device_cache_size="2G", cache_policy="device_replicate", csr_topo=topo)
    feature.from_cpu_tensor(x)
    batches = [b.to(dev) for b in bench.make_seed_batches(12)]
    n_id, _, _ = sampler.sample(batches[0])
    st = feature._my_store().shard_tensor
    outs = [torch.empty(n_id.numel() + 200000, bench.FEAT_DIM, device=dev) for _ in range(2)]
    big_idx = torch.randint(0, bench.N_NODES, (n_id.numel(), ), device=dev)
    for _ in range(3):
        st.gather(big_idx, feature.feature_order, out=outs[0][:big_idx.numel()])
        sampler.sample(batches[1])
    torch.cuda.synchronize()
    priv = torch.cuda.Stream(device=0, priority=-1)
    cur = torch.cuda.current_stream()
    res = {}

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def sample_on(stream, b):
        with torch.cuda.stream(stream):
            s0, s1 = ev(), ev()
            s0.record()
            sampler.quiver.sample_khop(b, bench.SIZES)
            s1.record()
        return s0, s1

    # 1. sampler alone on the private stream
    ts = []
    for b in batches[2:7]:
        s0, s1 = sample_on(priv, b)
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1))
    res["sampler_alone_ms"] = sorted(ts)[len(ts) // 2]
    # 2. gathers alone
    g0, g1 = ev(), ev()
    g0.record()
    for j in range(4):
        st.gather(big_idx, feature.feature_order, out=outs[j % 2][:big_idx.numel()])
    g1.record()
    torch.cuda.synchronize()
    res["four_gathers_alone_ms"] = g0.elapsed_time(g1)
    # 3. both: gathers on the current stream, sampler on the private one, enqueued right behind
    both = []
    for b in batches[7:12]:
        g0, g1 = ev(), ev()
        g0.record()
        for j in range(4):
            st.gather(big_idx, feature.feature_order, out=outs[j % 2][:big_idx.numel()])
        g1.record()
        s0, s1 = sample_on(priv, b)
        torch.cuda.synchronize()
        both.append({"gathers_ms": g0.elapsed_time(g1), "sampler_ms": s0.elapsed_time(s1),
                     "gather_start_to_sampler_end_ms": g0.elapsed_time(s1),
                     "gather_start_to_sampler_start_ms": g0.elapsed_time(s0)})
    res["concurrent"] = both
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
