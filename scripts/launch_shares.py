"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python scripts/launch_shares.py file.csv"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ix = {n: i for i, n in enumerate(rows[hdr])}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[hdr + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        tot[r[ix["Kernel Name"]]] += ms
        cnt[r[ix["Kernel Name"]]] += 1
    total = sum(tot.values())
    for k, v in tot.most_common():
        print("%-72s n=%4d total=%10.3f ms share=%5.1f%%" % (k[:72], cnt[k], v, 100 * v / total))


if __name__ == "__main__":
    main(sys.argv[1])
