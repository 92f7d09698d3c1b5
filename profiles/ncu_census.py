#!/usr/bin/env python
"""Regenerate profiles/kernel_census.json from an ncu --set full capture (dynamic counts, not estimates).

    python profiles/ncu_census.py <model> <capture.ncu-rep> <units_per_launch> <source_hash> [kernel-regex]

For the kernel of the capture: DRAM bytes per launch (dram__bytes_read + write), dynamic warp-instructions
per unit (a unit = a frame pair for CombSubFast, a frame for the filter models) and FMA-pipe slots per unit
(scalar FFMA / FMUL / FADD / IMAD one slot, packed FFMA2 / FMUL2 / FADD2 two) from the source page.
`source_hash` is the hash profiles/prof_stage.py printed in the run that was profiled (hash of csrc/);
bench.py only uses a record whose hash matches the sources it runs."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles', 'kernel_census.json')


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    model, rep, units, shash = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    pat = re.compile(sys.argv[5]) if len(sys.argv) > 5 else None
    raw = page(rep, 'raw')
    hdr, rows = raw[0], raw[2:]
    iN = hdr.index('Kernel Name')
    rows = [r for r in rows if pat is None or pat.search(r[iN])]
    kernels = []
    for r in rows:
        g = lambda k: float(r[hdr.index(k)])      # noqa: E731
        unit_r, unit_w = raw[1][hdr.index('dram__bytes_read.sum')], raw[1][hdr.index('dram__bytes_write.sum')]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        kernels.append({
            'dram_bytes': g('dram__bytes_read.sum') * scale[unit_r] + g('dram__bytes_write.sum') * scale[unit_w],
            'time_us': g('gpu__time_duration.sum'), 'inst': g('smsp__inst_executed.sum'),
            'issue_active_pct': g('smsp__issue_active.avg.pct_of_peak_sustained_active')})
    src = page(rep, 'source')
    blocks, cur, shdr, names = [], None, None, []
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = []
            blocks.append(cur)
            names.append(r[1])
        elif r and r[0] == 'Address':
            shdr = r
        elif cur is not None and len(r) > 5:
            cur.append(r)
    rec = {'source_hash': shash, 'capture': os.path.basename(rep), 'units_per_launch': units, 'kernels': {}}
    tot_dram = 0.0
    # the source page lists every profiled launch once per view (SASS, and again when source correlation is imported):
    # keep one block per launch so that they line up with the rows of the raw page
    if len(rows) and len(blocks) % len(raw[2:]) == 0 and len(blocks) > len(raw[2:]):
        step = len(blocks) // len(raw[2:])
        names, blocks = names[::step], blocks[::step]
    sel = [(n, b) for n, b in zip(names, blocks) if pat is None or pat.search(n)]
    for idx, (name, b) in enumerate(sel):
        iS, iE = shdr.index('Source'), shdr.index('Instructions Executed')
        slots = inst = packed = 0
        for r in b:
            m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS])
            op = m.group(1) if m else '?'
            n = int(r[iE])
            inst += n
            if op in ('FFMA2', 'FMUL2', 'FADD2'):
                slots += 2 * n
                packed += n
            elif op in ('FFMA', 'FMUL', 'FADD', 'IMAD'):
                slots += n
        short = re.sub(r'[<(].*', '', name).replace('void ', '').replace('ddsp::', '')
        k = kernels[idx] if idx < len(kernels) else {}      # raw page and source page list the launches in the same order
        cur = rec['kernels'].get(short)
        if cur is None:
            # dispatch cycles: a packed fp32x2 instruction holds the scheduler's dispatch port for two cycles
            # (profiles/ubench/ffma2_issue.cu), every other instruction for one
            rec['kernels'][short] = {'launches': 1, 'instructions_per_unit': inst / units, 'fma_pipe_slots_per_unit': slots / units,
                                     'dispatch_cycles_per_unit': (inst + packed) / units,
                                     'dram_bytes_per_launch': k.get('dram_bytes'), 'ncu_time_us': k.get('time_us'),
                                     'issue_active_pct': k.get('issue_active_pct')}
        else:       # the same kernel launched again inside the step (e.g. the two L = 510 convolutions): totals per step
            cur['launches'] += 1
            cur['instructions_per_unit'] += inst / units
            cur['fma_pipe_slots_per_unit'] += slots / units
            cur['dispatch_cycles_per_unit'] += (inst + packed) / units
            cur['dram_bytes_per_launch'] = (cur['dram_bytes_per_launch'] or 0.0) + (k.get('dram_bytes') or 0.0)
            cur['ncu_time_us'] = (cur['ncu_time_us'] or 0.0) + (k.get('time_us') or 0.0)
        tot_dram += k.get('dram_bytes') or 0.0
    rec['dram_bytes_per_launch'] = tot_dram
    rec['instructions_per_unit'] = sum(v['instructions_per_unit'] for v in rec['kernels'].values())
    rec['fma_pipe_slots_per_unit'] = sum(v['fma_pipe_slots_per_unit'] for v in rec['kernels'].values())
    rec['dispatch_cycles_per_unit'] = sum(v['dispatch_cycles_per_unit'] for v in rec['kernels'].values())
    rec['ncu_time_us'] = sum(v['ncu_time_us'] or 0.0 for v in rec['kernels'].values())
    if model == 'combsubfast' and rec['kernels']:
        k0 = next(iter(rec['kernels'].values()))
        rec['instructions_per_pair'] = k0['instructions_per_unit']
        rec['fma_pipe_slots_per_pair'] = k0['fma_pipe_slots_per_unit']
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    data[model] = rec
    with open(OUT, 'w') as f:
        json.dump(data, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == '__main__':
    main()
