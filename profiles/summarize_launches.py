"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python profiles/summarize_launches.py gpurun_out/<launches>.csv > profiles/<name>.md"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        t = float(r[vi].replace(",", ""))
        t = t / 1e3 if r[ui] == "ns" else (t * 1e3 if r[ui] == "ms" else t)
        name = r[ki]
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] = max(a[2], t)
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# ncu launch list summary: {path}\n")
    print(f"{n} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES, not absolutes)\n")
    print("| share | total us | launches | avg us | max us | kernel |")
    print("|---:|---:|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * v[1] / tot:.1f}% | {v[1]:.1f} | {v[0]} | {v[1] / v[0]:.1f} | {v[2]:.1f} | `{k[:150]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
