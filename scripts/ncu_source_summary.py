"""Aggregate an `ncu --page source --csv` export (SASS view): stall samples and instruction mix per kernel."""
import csv
import collections
import re
import sys


def main(path, top=25):
    kernels = []
    cur = None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1][:90], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None:
            cur["rows"].append(row)
    for k in kernels:
        h = k["hdr"]
        ix = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot = collections.Counter()
        mix = collections.Counter()
        samples = 0
        inst = 0
        for r in k["rows"]:
            try:
                ns = int(r[ix["# Samples"]] or 0)
                ni = int(r[ix["Instructions Executed"]] or 0)
            except ValueError:
                continue
            samples += ns
            inst += ni
            op = r[ix["Source"]].split()
            op = [t for t in op if not t.startswith("@")]
            mn = op[0].split(".")[0] if op else "?"
            mix[mn] += ni
            for s in stalls:
                try:
                    tot[s] += int(r[ix[s]] or 0)
                except ValueError:
                    pass
        print("==", k["name"])
        print("   samples", samples, "warp-instructions", inst)
        print("   stalls: " + ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(samples, 1)) for s, v in tot.most_common(8)))
        print("   mix: " + ", ".join("%s %.1f%%" % (m, 100.0 * v / max(inst, 1)) for m, v in mix.most_common(top)))


if __name__ == "__main__":
    main(sys.argv[1])
