#!/usr/bin/env python
"""Static SASS census of one kernel of libddsp_b200.so, by opcode and by source section.

    python profiles/sass_sections.py [--kernel combsubfast_kernelILb0] [--sections csf] [--json out.json]

Every SASS instruction is attributed to the OUTERMOST source line of its inline chain inside the
kernel's own file (nvdisasm -gi), and that line to a named section (line ranges given by the
`// @section name` markers in the source).  With the per-pair execution weights of each section this
gives warp-instructions and FMA-pipe slots per frame pair without a GPU; bench.py reads the JSON
this script writes (profiles/kernel_census.json) instead of hard-coded constants, and the build
hash recorded inside must match the library it runs.
"""
import argparse
import collections
import hashlib
import json
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'ddsp-svc-official_b200', 'lib', 'libddsp_b200.so')
CSRC = os.path.join(ROOT, 'ddsp-svc-official_b200', 'csrc')

# opcodes that occupy the FP32 (FMA) pipe: scalar ops one slot, packed fp32x2 two
FMA_SCALAR = ('FFMA', 'FMUL', 'FADD', 'IMAD', 'FSEL_NOT', )
FMA_PACKED = ('FFMA2', 'FMUL2', 'FADD2')


# the headers a model's synthesizer kernels are compiled from (a census record is tied to THESE files: edits to the
# control-network or front-end kernels do not invalidate it)
MODEL_SOURCES = {
    'combsubfast': ('common.cuh', 'fft32.cuh', 'phase.cuh', 'combsubfast.cuh'),
    'combsub': ('common.cuh', 'fft32.cuh', 'phase.cuh', 'excite.cuh', 'ltvfir.cuh'),
    'sins': ('common.cuh', 'fft32.cuh', 'phase.cuh', 'excite.cuh', 'ltvfir.cuh'),
}


def source_hash(model=None):
    """Hash of the kernel sources of `model` (all of csrc/ when None)."""
    h = hashlib.sha256()
    files = MODEL_SOURCES[model] if model else sorted(f for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh')))
    for f in files:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def disassemble(kernel_substr):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(['cuobjdump', '-xelf', 'all', LIB], cwd=td, check=True, capture_output=True)
        cubins = [f for f in os.listdir(td) if f.endswith('.cubin')]
        out = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(td, cubins[0])], capture_output=True, text=True).stdout
    lines = out.split('\n')
    start = None
    for i, l in enumerate(lines):
        if l.startswith('.text.') and kernel_substr in l:
            start = i
            break
    if start is None:
        raise SystemExit(f'kernel {kernel_substr} not found')
    end = len(lines)
    for i in range(start + 1, len(lines)):
        if lines[i].startswith('//---') or (lines[i].startswith('.text.') and kernel_substr not in lines[i]):
            end = i
            break
    return lines[start:end]


def section_map(path):
    """`// @section name` markers -> list of (first_line, name), sorted."""
    marks = []
    with open(path) as f:
        for n, l in enumerate(f, 1):
            m = re.search(r'//\s*@section\s+(\w+)', l)
            if m:
                marks.append((n, m.group(1)))
    return marks


def census(kernel_substr, kernel_file):
    dis = disassemble(kernel_substr)
    marks = section_map(kernel_file)
    base = os.path.basename(kernel_file)
    cur_outer, cur_inner = None, None
    chain = []
    per_section = collections.defaultdict(collections.Counter)
    per_line = collections.defaultdict(collections.Counter)
    pat_file = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    pat_ins = re.compile(r'/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
    pending = []
    for l in dis:
        m = pat_file.search(l)
        if m:
            pending.append((m.group(1), int(m.group(2)), m.group(3), int(m.group(4)) if m.group(4) else None))
            continue
        m = pat_ins.search(l)
        if not m:
            continue
        if pending:
            # chain: innermost first; the kernel-file line is the last entry that names the kernel file
            outer = None
            for (f, ln, f2, ln2) in pending:
                if os.path.basename(f) == base:
                    outer = ln
                if f2 and os.path.basename(f2) == base:
                    outer = ln2
            cur_inner = (os.path.basename(pending[0][0]), pending[0][1])
            if outer is not None:
                cur_outer = outer
            pending = []
        op = m.group(1)
        sec = 'prologue'
        if cur_outer is not None:
            for (n, name) in marks:
                if cur_outer >= n:
                    sec = name
        per_section[sec][op] += 1
        per_line[(sec, cur_inner)][op] += 1
    return per_section, per_line


def classify(op):
    root = op.split('.')[0]
    if root in FMA_PACKED:
        return 'fp32x2'
    if root in ('FFMA', 'FMUL', 'FADD'):
        return 'fp32'
    if root in ('DADD', 'DMUL', 'DFMA', 'DSETP', 'F2F', 'I2F', 'F2I', 'FRND', 'D2F'):
        return 'fp64/cvt'
    if root.startswith('MUFU'):
        return 'mufu'
    if root in ('LDS', 'STS', 'LDSM'):
        return 'smem'
    if root in ('LDG', 'STG', 'RED', 'ATOMG', 'LDL', 'STL', 'LD', 'ST', 'CCTL'):
        return 'gmem'
    if root in ('SHFL',):
        return 'shfl'
    if root in ('BRA', 'BSSY', 'BSYNC', 'EXIT', 'WARPSYNC', 'NANOSLEEP', 'BAR', 'CALL', 'RET', 'NOP', 'YIELD'):
        return 'ctrl'
    return 'int/other'


# executions of each section per frame pair of the CombSubFast kernel
CSF_WEIGHTS = {'excite': 2, 'frame': 2, 'fft': 3, 'filter': 2, 'stash': 1, 'pack': 1, 'ola': 1, 'loop': 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--kernel', default='combsubfast_kernelILb0')
    ap.add_argument('--file', default=os.path.join(CSRC, 'combsubfast.cuh'))
    ap.add_argument('--json', default=None)
    ap.add_argument('--lines', action='store_true', help='also list the hottest source lines')
    args = ap.parse_args()
    per_section, per_line = census(args.kernel, args.file)
    classes = ['fp32x2', 'fp32', 'fp64/cvt', 'mufu', 'smem', 'gmem', 'shfl', 'int/other', 'ctrl']
    print(f'{"section":10s} {"n":>6s} ' + ' '.join(f'{c:>9s}' for c in classes) + '   x/pair')
    tot_pair, fma_pair = 0, 0
    table = {}
    for sec, cnt in per_section.items():
        n = sum(cnt.values())
        by = collections.Counter()
        for op, k in cnt.items():
            by[classify(op)] += k
        w = CSF_WEIGHTS.get(sec, 0)
        slots = 2 * by['fp32x2'] + by['fp32'] + sum(k for op, k in cnt.items() if op.split('.')[0] == 'IMAD')
        tot_pair += w * n
        fma_pair += w * slots
        table[sec] = {'instructions': n, 'fma_pipe_slots': slots, 'weight_per_pair': w, **{c: by[c] for c in classes}}
        print(f'{sec:10s} {n:6d} ' + ' '.join(f'{by[c]:9d}' for c in classes) + f'   {w}')
    print(f'weighted per frame pair: {tot_pair} warp-instructions, {fma_pair} FMA-pipe slots')
    if args.lines:
        rows = sorted(per_line.items(), key=lambda kv: -sum(kv[1].values()))[:60]
        for (sec, inner), cnt in rows:
            top = ', '.join(f'{op}:{k}' for op, k in cnt.most_common(6))
            print(f'  {sec:8s} {inner[0]}:{inner[1]:<5d} {sum(cnt.values()):5d}  {top}')
    if args.json:
        with open(args.json, 'w') as f:
            json.dump({'kernel': args.kernel, 'source_hash': source_hash(), 'sections': table,
                       'instructions_per_pair': tot_pair, 'fma_pipe_slots_per_pair': fma_pair}, f, indent=1)


if __name__ == '__main__':
    main()
