"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.
usage: kernel_shares.py launches.csv out.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import OrderedDict


def main(src, dst, skip=0):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = OrderedDict()
    cnt = {}
    seen = 0
    for r in rows:
        if r is hdr or r[ki] == "Kernel Name":
            continue
        seen += 1
        if seen <= skip:
            continue
        name = re.sub(r"\(.*$", "", r[ki]).strip()
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        tot[name] = tot.get(name, 0.0) + v
        cnt[name] = cnt.get(name, 0) + 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_ms,share\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write('"%s",%d,%.4f,%.4f\n' % (k, cnt[k], v, v / total))
        f.write('"TOTAL",%d,%.4f,1.0\n' % (sum(cnt.values()), total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
