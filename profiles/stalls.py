"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest instructions.
usage: python profiles/stalls.py <source.csv> [top_n]"""
import csv
import sys


def main(path, topn=40):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in data)
    print(rows[0][1][:100])
    print("total samples", tot, " warp instructions", sum(int(r[ie]) for r in data))
    agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k:28s} {v:8d} {100 * v / tot:5.1f}%")
    print("--- hottest instructions (index, sass, samples, executed, top stall reasons)")
    top = sorted(enumerate(data), key=lambda kv: -int(kv[1][si]))[:topn]
    for idx, r in sorted(top):
        st = {hdr[i]: int(r[i]) for i in stalls if int(r[i]) > 0}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(idx, r[src].strip()[:64].ljust(64), r[si], r[ie], st)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
