"""Seeded synthetic graphs shared by the tests, the golden-vector generator, smoke() and bench.py."""
import numpy as np


def powerlaw_csr(n_nodes, mean_deg, seed=0, alpha=2.0, max_deg=None, zero_frac=0.02):
    """Reddit / products-shaped CSR: pareto(alpha) degrees scaled to `mean_deg` (a few isolated nodes), uniform
    neighbours, columns sorted inside each row (what scipy's COO->CSR gives the reference: SURVEY.md 8(a1))."""
    rng = np.random.default_rng(seed)
    raw = rng.pareto(alpha, n_nodes) + 1.0
    deg = np.floor(raw * (mean_deg / raw.mean())).astype(np.int64)
    deg = np.minimum(deg, (max_deg if max_deg is not None else n_nodes - 1))
    deg[rng.random(n_nodes) < zero_frac] = 0
    indptr = np.zeros(n_nodes + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = rng.integers(0, n_nodes, int(indptr[-1]), dtype=np.int64)
    # sort columns within each row
    row = np.repeat(np.arange(n_nodes, dtype=np.int64), deg)
    order = np.lexsort((indices, row))
    return indptr, indices[order]


def rmat_csr(scale, n_edges, n_nodes=None, abcd=(0.57, 0.19, 0.19, 0.05), seed=2, dedup=False):
    """BASELINE config 2 shape: R-MAT edges with (a,b,c,d) = (0.57,0.19,0.19,0.05) over 2^scale ids,
    trimmed to `n_nodes`; duplicates kept unless `dedup` (a multigraph row is legal CSR input: sampling is by POSITION).
    Heavy skew, many empty rows, duplicate neighbours -- a different degree structure from the pareto graphs."""
    rng = np.random.default_rng(seed)
    a, b, c, _ = abcd
    src = np.zeros(n_edges, np.int64)
    dst = np.zeros(n_edges, np.int64)
    for _level in range(scale):
        r = rng.random(n_edges)
        down = r >= a + b  # quadrants c, d: lower half (src bit set)
        right = ((r >= a) & (r < a + b)) | (r >= a + b + c)  # quadrants b, d: right half (dst bit set)
        src = (src << 1) | down
        dst = (dst << 1) | right
    n = n_nodes if n_nodes is not None else (1 << scale)
    keep = (src < n) & (dst < n)
    src, dst = src[keep], dst[keep]
    if dedup:
        key = np.unique(src * n + dst)
        src, dst = key // n, key % n
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=indptr[1:])
    return indptr, dst


def simple_graph(n, nbr):
    """The reference's own sampler fixture: node i has neighbours (j+1)*n + i (tests/cpp/test_quiver_cpu.cpp:9-30).
    Neighbour ids exceed n, so the CSR is padded with empty rows up to the largest id."""
    total = (nbr + 1) * n
    deg = np.zeros(total, np.int64)
    deg[:n] = nbr
    indptr = np.zeros(total + 1, np.int64)
    np.cumsum(deg, out=indptr[1:])
    indices = (np.arange(1, nbr + 1, dtype=np.int64)[None, :] * n + np.arange(n, dtype=np.int64)[:, None]).reshape(-1)
    return indptr, indices


MINI = dict(  # known-answer fixture produced by the reference CPU build
    indptr=[0, 3, 5, 5, 9], indices=[1, 2, 3, 0, 2, 0, 1, 2, 3], seeds=[0, 3, 2, 1], k=2,
    counts=[2, 2, 0, 2], draw=[1, 3, 1, 3, 0, 2],
    frontier=[0, 3, 2, 1], row_idx=[0, 0, 1, 1, 3, 3], col_idx=[3, 1, 3, 1, 0, 2])
