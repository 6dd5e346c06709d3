"""Reads the timeline the tcgen05 eval kernel writes with B2E_TC_TRACE=<file> (CTA 0, last launch) and
prints the latency chain of the forward ring: issue -> converted -> MMA issued -> stage refilled."""
import sys

import numpy as np

roles = {}
for line in open(sys.argv[1]):
    parts = line.split()
    roles[parts[0]] = np.array([int(v) for v in parts[1:]], dtype=np.int64)
t0 = min(v[0] for v in roles.values() if len(v))
for k in roles:
    roles[k] = roles[k] - t0
SF = 4
n = min(len(roles['w_issue']), len(roles['xf_issue']), len(roles['conv_f']), len(roles['mma_f']))
w, x, c, m = (roles[k][:n] for k in ('w_issue', 'xf_issue', 'conv_f', 'mma_f'))
print('units traced', n, '| cycles; 1 us = ~1900 cycles')
print('cadence mma_f[n+1]-mma_f[n]: median %d mean %d' % (np.median(np.diff(m)), np.mean(np.diff(m))))
print('load+convert  conv_f - max(w_issue, xf_issue): median %d' % np.median(c - np.maximum(w, x)))
print('xf_issue - w_issue: median %d' % np.median(x - w))
print('MMA pickup    mma_f - conv_f: median %d' % np.median(m - c))
print('refill        w_issue[n+%d] - mma_f[n]: median %d' % (SF, np.median(w[SF:] - m[:-SF])))
print('first 60 units: w_issue xf_issue conv_f mma_f')
for i in range(min(n, 60)):
    print(i, w[i], x[i], c[i], m[i])
for k in ('mma_b', 'xb_issue', 'tail', 'drain'):
    print(k, roles[k][:40].tolist())
