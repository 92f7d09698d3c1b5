#!/usr/bin/env python
"""Dynamic instruction census of combsubfast_kernel per source section (`// @section name` markers): the per-SASS
instruction execution counts of an ncu capture (--set full --import-source on) joined, in program order, with the
section attribution of profiles/sass_sections.py for the library the capture profiled.

    python profiles/ncu_sections.py gpurun_out/r02_census_combsubfast.ncu-rep [units_per_launch=27648]

Columns: warp-instructions and dispatch cycles (packed fp32x2 counted twice, see profiles/ubench/ffma2_issue.cu) per
frame pair, then instruction classes.  The library in ddsp-svc-official_b200/lib must be the build that was profiled."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from profiles import sass_sections as S  # noqa: E402


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 27648.0
    kernel, kfile = 'combsubfast_kernelILb0', os.path.join(S.CSRC, 'combsubfast.cuh')
    dis, marks, base = S.disassemble(kernel), S.section_map(kfile), os.path.basename(kfile)
    pat_file = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    pat_ins = re.compile(r'/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
    seq, pending, cur_outer = [], [], None
    for l in dis:
        m = pat_file.search(l)
        if m:
            pending.append((m.group(1), int(m.group(2)), m.group(3), int(m.group(4)) if m.group(4) else None))
            continue
        m = pat_ins.search(l)
        if not m:
            continue
        if pending:
            outer = None
            for (f, ln, f2, ln2) in pending:
                if os.path.basename(f) == base:
                    outer = ln
                if f2 and os.path.basename(f2) == base:
                    outer = ln2
            if outer is not None:
                cur_outer = outer
            pending = []
        sec = 'prologue'
        if cur_outer is not None:
            for (n, name) in marks:
                if cur_outer >= n:
                    sec = name
        seq.append((m.group(2), sec))
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    hdr, dyn = None, []
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == 'Address':
            hdr = r
            iE = hdr.index('Instructions Executed')
        elif hdr and len(r) > 5 and r[0].startswith('0x'):
            dyn.append(int(r[iE]) if r[iE].isdigit() else 0)
    if len(seq) != len(dyn):
        raise SystemExit(f'library ({len(seq)} SASS instructions) is not the build the capture profiled ({len(dyn)})')
    agg, ops = collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter)
    for (op, sec), cnt in zip(seq, dyn):
        root = op.split('.')[0]
        agg[sec][S.classify(op)] += cnt
        agg[sec]['_n'] += cnt
        agg[sec]['_disp'] += (2 if root in S.FMA_PACKED else 1) * cnt
        ops[sec][root] += cnt
    classes = ['fp32x2', 'fp32', 'fp64/cvt', 'mufu', 'smem', 'gmem', 'shfl', 'int/other', 'ctrl']
    print(f'{"section":10s} {"instr":>8s} {"dispatch":>8s} ' + ' '.join(f'{c:>9s}' for c in classes))
    tot = collections.Counter()
    for sec, c in sorted(agg.items(), key=lambda kv: -kv[1]['_disp']):
        print(f'{sec:10s} {c["_n"] / units:8.1f} {c["_disp"] / units:8.1f} ' + ' '.join(f'{c[k] / units:9.1f}' for k in classes))
        tot.update(c)
    print(f'{"total":10s} {tot["_n"] / units:8.1f} {tot["_disp"] / units:8.1f} ' + ' '.join(f'{tot[k] / units:9.1f}' for k in classes))
    print()
    for sec, c in sorted(ops.items(), key=lambda kv: -sum(kv[1].values())):
        print(f'{sec:10s}', ', '.join(f'{k}:{v / units:.0f}' for k, v in c.most_common(16) if v / units >= 0.5))


if __name__ == '__main__':
    main()
