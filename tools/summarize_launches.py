#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel body (count, total, average, share)."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        m = re.search(r"<(?:pcgpu::)?(\w+)(?:<pcgpu::(\w+)(?:, \(bool\)(\d))?)?", r[kn])
        key = r[kn][:50]
        if m:
            key = m.group(1)
            if m.group(2):
                key += f"<{m.group(2)}{',r0' if m.group(3) == '1' else (',r>0' if m.group(3) == '0' else '')}>"
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    for k, (c, v) in agg.items():
        print(f"{k:45s} n={c:4d} total={v / 1e6:9.3f} ms  avg={v / c / 1e3:9.1f} us  share={100 * v / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
