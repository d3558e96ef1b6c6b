"""Aggregate an `ncu --import-source on` capture per CUDA source line.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > both.csv
       python profiles/ncu_source_lines.py both.csv [min_percent]
Prints, for every line holding more than min_percent of the stall samples or of the executed
warp instructions: share of samples, share of instructions.
"""
import collections
import csv
import sys


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(path, min_pct=0.8):
    rows = list(csv.reader(open(path)))
    k = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr, data = rows[k], rows[k + 1:]
    isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
    agg, cur = collections.OrderedDict(), None
    for r in data:
        if len(r) <= iex:
            continue
        if r[0] != "":
            cur = (r[0], r[1])
            agg.setdefault(cur, [0, 0])
        if r[2] != "" and cur:
            agg[cur][0] += num(r[isamp])
            agg[cur][1] += num(r[iex])
    ts = sum(v[0] for v in agg.values()) or 1
    te = sum(v[1] for v in agg.values()) or 1
    print("samples %d  warp instructions %d" % (ts, te))
    for (ln, src), (s, e) in agg.items():
        if s > min_pct / 100 * ts or e > min_pct / 100 * te:
            print("%5s %-95s smp %5.1f%%  ex %5.1f%%" % (ln, src.strip()[:95], 100 * s / ts, 100 * e / te))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.8)
