#!/usr/bin/env python
"""Per source line dynamic instruction counts and stall samples of one kernel in an .ncu-rep
(captured with --import-source on, built with -lineinfo).
usage: python profiles/ncu_lines.py rep.ncu-rep [units_per_launch] [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, hdr, recs = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fpath = r[1].split('/')[-1]
        elif r[0] == 'Line No':
            hdr = r
            iI = hdr.index('Instructions Executed')
            iS = hdr.index('Warp Stall Sampling (All Samples)')
            iN = hdr.index('Warp Stall Sampling (Not-issued Samples)')
        elif hdr and r[0].isdigit():
            f = lambda v: int(v) if v.lstrip('-').isdigit() else 0
            recs.append((fpath, int(r[0]), r[1].strip(), f(r[iI]), f(r[iS]), f(r[iN])))
    totI = sum(x[3] for x in recs)
    totS = sum(x[4] for x in recs)
    print(f'total {totI} warp-instructions = {totI / units:.1f}/unit, {totS} stall samples')
    byfile = {}
    for f, ln, src, i, s, n in recs:
        a = byfile.setdefault(f, [0, 0, 0])
        a[0] += i; a[1] += s; a[2] += n
    for f, (i, s, n) in byfile.items():
        print(f'{f:24s} instr {i / units:8.1f}/unit {100 * i / totI:5.1f}%  samples {100 * s / totS:5.1f}%  not-issued {100 * n / totS:5.1f}%')
    print()
    recs.sort(key=lambda x: -x[4])
    for f, ln, src, i, s, n in recs[:top]:
        print(f'{f:18s}:{ln:4d} instr {i / units:7.1f} ({100 * i / totI:4.1f}%) samples {100 * s / totS:4.1f}% cyc/instr {s / max(i, 1) * totI / totS:5.2f} | {src[:90]}')


if __name__ == '__main__':
    main()
