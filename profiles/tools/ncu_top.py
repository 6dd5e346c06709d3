"""Summarise an ncu report of one kernel: headline counters plus the SASS lines with the most stall samples.
usage: python profiles/tools/ncu_top.py report.ncu-rep [kernel-index] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(' | '.join('%s=%s%s' % (k.split('.')[0][-28:], d[k], units[hdr.index(k)][:6]) for k in KEYS if k in d))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}
        kern.append(cur)
    elif r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and r:
        cur['rows'].append(r)
k = kern[which]
h = k['hdr']
i_s, i_src, i_ex = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
total = sum(int(r[i_s]) for r in k['rows'])
print(k['name'][:80], 'samples', total, 'warp instructions', sum(int(r[i_ex]) for r in k['rows']))
top = sorted(range(len(k['rows'])), key=lambda i: -int(k['rows'][i][i_s]))[:topn]
for i in sorted(top):
    r = k['rows'][i]
    st = sorted(((int(r[c]) if r[c] else 0, h[c]) for c in stall), reverse=True)[:2]
    print('%5d %6s %9s  %-70s %s' % (i, r[i_s], r[i_ex], r[i_src].strip()[:70], st))
