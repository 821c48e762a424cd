"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share, and the last
step's launch sequence.  Usage: python profiles/launch_table.py gpurun_out/launches.csv [n_tail]"""
import collections
import csv
import sys


def main(path, tail=24):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, seq = collections.OrderedDict(), []
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].split("::")[-1]
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        seq.append((name, v))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':44s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {a[0]:5d} {a[1]:10.1f} {a[1] / a[0]:9.2f} {a[1] / tot:7.1%}")
    print(f"-- last {tail} launches --")
    for n, v in seq[-tail:]:
        print(f"   {n:40s} {v:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
