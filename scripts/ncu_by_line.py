#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i rep --page source --csv --kernel-id ...)
with nvdisasm line info of the cubin, and aggregate executed instructions and
stall samples per CUDA source line.

usage: ncu_by_line.py <sass.csv> <cubin> <kernel-substring> [top]
"""
import collections
import csv
import re
import subprocess
import sys


def line_map(cubin, kernel):
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True,
                         text=True).stdout
    lines = txt.splitlines()
    out, cur, active = [], None, False
    for ln in lines:
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            active = kernel in m.group(1)
            continue
        if ln.startswith('\t.section') or ln.startswith('.section'):
            active = False
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
            op = ln.split('*/', 1)[1].strip().rstrip(';')
            out.append((cur, op))
    return out


def main():
    sass_csv, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    insts = [dict(zip(hdr, r)) for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    lm = line_map(cubin, kernel)
    if len(insts) > len(lm) and len(insts) % len(lm) == 0:
        insts = insts[:len(lm)]      # several launches captured: use the first
    if len(lm) != len(insts):
        print(f'warning: {len(insts)} profiled vs {len(lm)} disassembled')
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for (loc, _), d in zip(lm, insts):
        n = int(d['Instructions Executed'] or 0)
        s = int(d['# Samples'] or 0)
        agg[loc][0] += n
        agg[loc][1] += s
        agg[loc][2] += 1
        tot_i += n
        tot_s += s
    print(f'total warp-instructions {tot_i}, samples {tot_s}')
    src = {}
    for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ''
        if loc:
            try:
                if loc[0] not in src:
                    import glob
                    path = glob.glob(f'/root/repo/pointvs_b200/csrc/{loc[0]}')
                    src[loc[0]] = open(path[0]).read().splitlines() if path else []
                text = src[loc[0]][loc[1] - 1].strip()[:80]
            except (IndexError, OSError):
                pass
        print(f'{str(loc):32s} sass={v[2]:4d} inst={v[0] / max(1, tot_i):6.1%} '
              f'samples={v[1] / max(1, tot_s):6.1%} | {text}')


if __name__ == '__main__':
    main()
