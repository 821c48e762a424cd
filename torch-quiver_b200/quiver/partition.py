"""Access-probability feature partitioning (the placement `north_star` assumes).
Reference: srcs/python/quiver/partition.py:16-283.  Offline, once per graph: `GraphSageSampler.sample_prob` (the
`cal_next` kernel behind `Quiver.cal_neighbor_prob`) gives every trainer's per-node access probability; nodes are then
dealt to partitions so that each partition holds the nodes IT reads most and the others read least.

Same greedy as the reference, expressed on a [P, blob] score matrix instead of P nested Python loops:
  blob     = `chunk_size * P` consecutive node ids
  score_r  = P * probs[r] - sum_{q != r} probs[q]  (+1e-6, accumulated in the reference's order)
  partition r (round-robin start, rotating one position per blob) takes its `chunk_size` best-scoring unclaimed nodes.
"""
import os
import shutil
from typing import List, Optional

import torch

from . import utils as quiver_util

__all__ = ["quiver_partition_feature", "load_quiver_feature_partition", "partition_without_replication",
           "partition_feature_without_replication", "select_nodes"]

QUIVER_MAGIC_NUMBER = 256
CHUNK_NUM = 32


def _default_device():
    return torch.cuda.current_device() if torch.cuda.is_available() else "cpu"


def _scores(probs, chunk):
    """[P, len(chunk)] affinity of every partition for every node of the chunk (reference: partition.py:51-60)."""
    P = len(probs)
    rows = []
    for r in range(P):
        sc = torch.zeros(chunk.numel(), device=chunk.device) + 1e-6
        for q in range(P):  # same accumulation order as the reference, so ties break identically
            if q == r:
                sc += probs[q][chunk] * P
            else:
                sc -= probs[q][chunk]
        rows.append(sc)
    return torch.stack(rows)


def _deal(scores, start_rank, per_part, taken_value):
    """Round-robin greedy over one blob.  Returns {rank: positions inside the blob}."""
    P, n = scores.shape
    picks, assigned = {}, 0
    for r_ in range(start_rank, start_rank + P):
        r = r_ % P
        size = min(per_part, n - assigned)
        order = torch.sort(scores[r], descending=True)[1][:size]
        picks[r] = order
        scores[:, order] = taken_value  # nobody else may claim them
        assigned += size
    return picks


def partition_without_replication(device, probs, ids: Optional[torch.Tensor]):
    """Split `ids` (or all nodes) into len(probs) disjoint parts by access probability (partition.py:16-84):
    CHUNK_NUM equal chunks, each dealt evenly."""
    P = len(probs)
    ids = ids.to(device) if ids is not None else None
    probs = [(p[ids] if ids is not None else p).to(device) for p in probs]
    total = probs[0].size(0)
    res = [[] for _ in range(P)]
    chunk_size = (total + CHUNK_NUM - 1) // CHUNK_NUM
    beg = 0
    for i in range(CHUNK_NUM):
        end = min(total, beg + chunk_size)
        if end <= beg:
            break
        chunk = torch.arange(beg, end, dtype=torch.int64, device=device)
        picks = _deal(_scores(probs, chunk), i, (chunk.numel() + P - 1) // P, -1e6)
        for r, pos in picks.items():
            res[r].append(chunk[pos])
        beg = end
    out = []
    for r in range(P):
        part = torch.cat(res[r]) if res[r] else torch.empty(0, dtype=torch.int64, device=device)
        out.append(ids[part] if ids is not None else part)
    return out


def select_nodes(device, probs, ids):
    """Summed access probability and the nodes anyone touches (partition.py:87-96)."""
    prob_sum = torch.zeros(probs[0].size(0), device=device)
    for prob in probs:
        if ids is None:
            prob_sum += prob.to(device)
        else:
            prob_sum[ids] += prob.to(device)[ids]
    return prob_sum, torch.nonzero(prob_sum)


def partition_feature_without_replication(probs: List[torch.Tensor], chunk_size: int, device=None):
    """Partition ALL nodes: blobs of chunk_size*P consecutive ids, every partition takes chunk_size of each blob
    (partition.py:99-161).  Returns (list of id tensors, probs moved to the device)."""
    device = _default_device() if device is None else device
    P = len(probs)
    probs = [p.to(device) for p in probs]
    total = probs[0].size(0)
    res = [[] for _ in range(P)]
    blob = chunk_size * P
    start, rot = 0, 0
    while start < total:
        end = min(total, start + blob)
        chunk = torch.arange(start, end, device=device)
        picks = _deal(_scores(probs, chunk), rot, chunk_size, -1.0)
        for r, pos in picks.items():
            res[r].append(chunk[pos])
        rot += 1
        start = end
    return [torch.cat(r) if r else torch.empty(0, dtype=torch.int64, device=device) for r in res], probs


def quiver_partition_feature(probs, result_path: str, cache_memory_budget=0, per_feature_size=0,
                             chunk_size=QUIVER_MAGIC_NUMBER, overwrite: bool = False, device=None):
    """Partition + per-partition hot set, written as the reference's folder layout (partition.py:164-249):

        result_path/feature_partition_{i}/partition_res.pth, cache_res.pth ; result_path/feature_partition_book.pth

    The reference prompts on stdin when `result_path` exists; here `overwrite=True` replaces it, otherwise
    FileExistsError.  Returns (partition_book, partition_res, cache_res)."""
    device = _default_device() if device is None else device
    if os.path.exists(result_path):
        if not overwrite:
            raise FileExistsError(f"{result_path} already exists (pass overwrite=True to replace it)")
        shutil.rmtree(result_path)
    P = len(probs)
    for i in range(P):
        os.makedirs(os.path.join(result_path, f"feature_partition_{i}"))
    cache_count = int(quiver_util.parse_size(cache_memory_budget) / (quiver_util.parse_size(per_feature_size) + 1e-6))
    per_partition_cache = cache_count // P
    partition_res, moved = partition_feature_without_replication(probs, chunk_size, device)
    partition_book = torch.zeros(moved[0].shape, dtype=torch.int64, device=device)
    cache_res = [None] * P
    for i in range(P):
        if cache_count > 0:
            cache_res[i] = torch.sort(moved[i], descending=True)[1][:per_partition_cache]
        partition_book[partition_res[i]] = i
        torch.save(partition_res[i], os.path.join(result_path, f"feature_partition_{i}", "partition_res.pth"))
        torch.save(cache_res[i], os.path.join(result_path, f"feature_partition_{i}", "cache_res.pth"))
    torch.save(partition_book, os.path.join(result_path, "feature_partition_book.pth"))
    return partition_book, partition_res, cache_res


def load_quiver_feature_partition(partition_idx: int, result_path: str):
    """(partition_book, this partition's ids, its cached ids) -- partition.py:252-283."""
    if not os.path.exists(result_path):
        raise Exception("Result path not exists")
    d = os.path.join(result_path, f"feature_partition_{partition_idx}")
    return (torch.load(os.path.join(result_path, "feature_partition_book.pth")),
            torch.load(os.path.join(d, "partition_res.pth")), torch.load(os.path.join(d, "cache_res.pth")))
