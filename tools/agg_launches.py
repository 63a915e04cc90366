"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/agg_launches.py launches.csv [start_marker_occurrence]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    occ = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]) for r in csv.DictReader(lines)]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[rows[0][2]]
    names = [r[0] for r in rows]
    marks = [i for i, n in enumerate(names) if "qsample_mix" in n]
    start = marks[occ] if marks else 0
    seg = rows[start:]
    tot = sum(r[1] for r in seg) * scale
    print(f"launches in the last training step: {len(seg)}, serialized kernel time {tot:.3f} ms")
    agg = collections.OrderedDict()
    for n, v, _ in seg:
        k = re.sub(r"\(.*", "", n).replace("sinddm::", "").replace("<unnamed>::", "").replace("void ", "")[:64]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v:9.3f} ms {c:4d}x {100 * v / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main()
