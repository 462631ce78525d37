"""Summarise a multi-metric `ncu --metrics ... --csv` launch list per kernel (time share, instructions, DRAM bytes).
usage: python profiles/launches2.py gpurun_out/<launches>.csv [n_updates]"""
import collections
import csv
import sys


def main(path, nup=2):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = {}
    for r in data:
        if len(r) <= vi:
            continue
        d = per.setdefault(int(r[0]), {"name": r[ki]})
        d[r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.OrderedDict()
    for i, d in sorted(per.items()):
        n = d["name"].split("(")[0].replace("void ", "").replace("dgvit::", "")
        a = agg.setdefault(n, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        t = d["gpu__time_duration.sum"] / 1e3
        a[0] += 1; a[1] += t; a[2] = max(a[2], t)
        a[3] += d.get("smsp__inst_executed.sum", 0)
        a[4] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        a[5] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * t
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: {path}\n\n{len(per)} launches over {nup} update(s), {tot:.1f} us total "
          f"(cold-cache, serialised: compare SHARES, not absolutes)\n")
    print("| share | us / update | launches / update | avg us | max us | warp instr / launch | DRAM MB / launch | tensor pipe active % | kernel |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * a[1] / tot:.1f}% | {a[1] / nup:.1f} | {a[0] / nup:g} | {a[1] / a[0]:.1f} | {a[2]:.1f} | {a[3] / a[0] / 1e6:.2f} M | "
              f"{a[4] / a[0] / 1e6:.1f} | {a[5] / a[1]:.1f} | `{n[:90]}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
