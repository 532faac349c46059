#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel shares.

    python tools/launch_shares.py gpurun_out/launches.csv [out.txt]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for i, r in enumerate(rows):
        if r and r[0] == 'ID':
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] == 'ns' else v * 1000 if r[ui] == 'ms' else v
        agg[r[ki][:100]][0] += 1
        agg[r[ki][:100]][1] += v
        tot += v
    lines = [f"# total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache serialised ncu times: compare SHARES)"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{t:10.1f} us {100 * t / tot:5.1f}% n={n:4d} {k}")
    text = "\n".join(lines)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + "\n")
    print("\n".join(lines[:28]))


if __name__ == '__main__':
    main()
