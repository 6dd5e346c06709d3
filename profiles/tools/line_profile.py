"""Aggregate an ncu SASS-level source page by CUDA source line.

usage: line_profile.py <report.ncu-rep> <library.so> [kernel-substring] [top-N]
Needs ncu, cuobjdump and nvdisasm on PATH (works without a GPU)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so_path, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so_path)], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    text = subprocess.run(['nvdisasm', '-g', '-c', cubin], check=True, capture_output=True, text=True).stdout
    lines, inside, cur = [], False, None
    for row in text.splitlines():
        if row.startswith('//---') and '.text.' in row:
            inside = kernel in row
            continue
        if not inside:
            continue
        m = re.search(r'//## File ".*?", line (\d+)(?: inlined at ".*?", line (\d+))?', row)
        if m:
            cur = int(m.group(1))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', row):
            lines.append(cur)
    return lines


def main():
    rep, so_path = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else 'optenv_kernel'
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(['ncu', '-i', rep, '--kernel-name', 'regex:' + kernel.split('I')[0].split('(')[0],
                          '--page', 'source', '--csv'], check=True,
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    col = {name: hdr.index(name) for name in ('Source', '# Samples', 'Instructions Executed',
                                             'L1 Wavefronts Shared', 'L2 Theoretical Sectors Global')}
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr) and r[0] != 'Address']
    lines = sass_lines(so_path, kernel)
    if len(lines) != len(body):
        print('warning: %d SASS rows in the report vs %d in the library' % (len(body), len(lines)))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    total = [0, 0, 0, 0]
    for line, row in zip(lines, body):
        vals = [int(float(row[col[k]] or 0)) for k in ('# Samples', 'Instructions Executed',
                                                        'L1 Wavefronts Shared', 'L2 Theoretical Sectors Global')]
        for i, v in enumerate(vals):
            agg[line][i] += v
            total[i] += v
    src = open(os.path.join(os.path.dirname(so_path), 'csrc', 'b200env.cu')).read().splitlines() \
        if os.path.exists(os.path.join(os.path.dirname(so_path), 'csrc', 'b200env.cu')) else []
    print('total samples %d, warp instructions %d, smem wavefronts %d, L2 sectors %d' % tuple(total))
    print('%6s %7s %7s %7s %7s  %s' % ('line', 'samp%', 'inst%', 'smem%', 'l2sec%', 'source'))
    for line, vals in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src[line - 1].strip()[:90] if line and line <= len(src) else ''
        print('%6s %7.2f %7.2f %7.2f %7.2f  %s' % (line, 100 * vals[0] / max(total[0], 1),
                                                   100 * vals[1] / max(total[1], 1),
                                                   100 * vals[2] / max(total[2], 1),
                                                   100 * vals[3] / max(total[3], 1), text))


if __name__ == '__main__':
    main()
