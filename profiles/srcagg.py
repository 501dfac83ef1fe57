"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.

    python profiles/srcagg.py <export.csv> [top_n]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    cur = None
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0])
    srcs = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or cur is None:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        def num(name):
            try:
                return int(r[hdr.index(name)])
            except (ValueError, IndexError):
                return 0
        agg[(cur, ln)][0] += num("Instructions Executed")
        agg[(cur, ln)][1] += num("# Samples")
        srcs[(cur, ln)] = r[1]
    tot = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    print("warp instructions", tot, "samples", ts)
    byfile = collections.Counter()
    for (f, _), v in agg.items():
        byfile[f] += v[0]
    print({k: f"{100 * v / tot:.1f}%" for k, v in byfile.items()})
    for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{f}:{l:4d} inst {100 * v[0] / tot:5.2f}% samp {100 * v[1] / max(ts, 1):5.2f}%  {srcs[(f, l)].strip()[:110]}")


if __name__ == "__main__":
    main()
