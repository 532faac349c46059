#!/usr/bin/env python
"""Per-instruction warp-stall samples of an .ncu-rep (needs --import-source on / -lineinfo).

    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [kernel-regex] [launch-index] [top-n]
"""
import csv
import subprocess
import sys

COLS = ['stall_barrier', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_short_sb', 'stall_wait',
        'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_lg', 'stall_branch_resolving',
        'stall_membar', 'stall_no_inst', 'stall_sleep', 'stall_drain', 'stall_tex', 'stall_misc']


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else '.'
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    hdr = rows[secs[which] + 1]
    body = rows[secs[which] + 2:secs[which + 1]]
    si, src = hdr.index('# Samples'), hdr.index('Source')
    ci = [(c, hdr.index(c)) for c in COLS if c in hdr]
    body = [r for r in body if len(r) > si and r[si].isdigit()]
    tot = sum(int(r[si]) for r in body)
    agg = {c: sum(int(r[i]) for r in body if r[i].isdigit()) for c, i in ci}
    print(rows[secs[which]][1][:100])
    print('total samples', tot)
    print({k: f'{100 * v / tot:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    for r in sorted(body, key=lambda r: -int(r[si]))[:topn]:
        st = {c[6:]: r[i] for c, i in ci if r[i] not in ('0', '')}
        print(f'{int(r[si]):6d} {r[src][:70]:70s} {st}')


if __name__ == '__main__':
    main()
