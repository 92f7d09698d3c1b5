#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the few numbers DESIGN.md / bench.py quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [units_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = page(rep, 'raw')
    hdr, unit, data = raw[0], raw[1], raw[2:]
    keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
            'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sm__cycles_elapsed.max',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'launch__grid_size', 'launch__block_size']
    for row in data:
        print('---')
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                print(f'{k:70s} {row[i]} {unit[i]}')
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
                v = float(row[i])
                if v >= 0.1:
                    print(f'  stall {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:28s} {v:.3f}')
    src = page(rep, 'source')
    blocks, cur, shdr = [], None, None
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = []
            blocks.append(cur)
        elif r and r[0] == 'Address':
            shdr = r
        elif cur is not None and len(r) > 5:
            cur.append(r)
    if blocks and shdr:
        b = blocks[0]
        iS, iE, iT = shdr.index('Source'), shdr.index('Instructions Executed'), shdr.index('Warp Stall Sampling (All Samples)')
        tot = sum(int(r[iE]) for r in b)
        ops, st = collections.Counter(), collections.Counter()
        for r in b:
            m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS])
            op = m.group(1) if m else '?'
            ops[op] += int(r[iE])
            st[op] += int(r[iT])
        S = max(1, sum(st.values()))
        print(f'--- static SASS instructions {len(b)}, dynamic warp instructions {tot}')
        for k, v in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 24):
            per = f'{v / units:9.1f}/unit' if units else ''
            print(f'{k:10s} {100 * v / tot:5.1f}% {per}  stall-samples {100 * st[k] / S:5.1f}%')


if __name__ == '__main__':
    main()
