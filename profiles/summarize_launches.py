"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel the launch
count, mean device time and share of the total.  (Per-launch times under ncu are cold-cache and
serialised: compare SHARES, not absolutes.)   python profiles/summarize_launches.py file.csv"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("nrc::", "")[:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else v          # -> us
        agg.setdefault(name, []).append(v)
    total = sum(sum(v) for v in agg.values())
    print("%-70s %6s %10s %8s" % ("kernel", "n", "mean us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-70s %6d %10.2f %7.1f%%" % (k, len(v), sum(v) / len(v), 100.0 * sum(v) / total))


if __name__ == "__main__":
    main(sys.argv[1])
