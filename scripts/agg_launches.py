"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python scripts/agg_launches.py file.csv [skip_until_substring [occurrence]]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    seq = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        seq.append((name, v))
    if len(sys.argv) > 2:
        occ = int(sys.argv[3]) if len(sys.argv) > 3 else 1
        idx = [i for i, (n, _) in enumerate(seq) if sys.argv[2] in n]
        seq = seq[idx[occ - 1] + 1:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in seq:
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v for _, v in seq)
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-64s %6d %11.1f us %5.1f%%  avg %9.1f us" % (n[:64], c, v, 100 * v / tot, v / c))
    print("total %.3f ms over %d launches" % (tot / 1e3, len(seq)))


if __name__ == "__main__":
    main()
