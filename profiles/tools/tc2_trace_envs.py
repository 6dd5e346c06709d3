"""Per-env view of the B2E_TC_TRACE timeline: forward-phase length, gap to the next env, tail duration."""
import sys

import numpy as np

roles = {}
for line in open(sys.argv[1]):
    p = line.split()
    roles[p[0]] = np.array([int(v) for v in p[1:]], dtype=np.int64)
t0 = min(v[0] for v in roles.values() if len(v))
m, w, c = (roles[k] - t0 for k in ('mma_f', 'w_issue', 'conv_f'))
x = roles['w_landed'] - t0 if len(roles.get('w_landed', [])) >= len(w) else w
tail, mb, dr = roles['tail'] - t0, roles['mma_b'] - t0, roles['drain'] - t0
UF, TB = 25, 7
for k in range(min(len(m) // UF, int(sys.argv[2]) if len(sys.argv) > 2 else 12)):
    a, b = UF * k, UF * k + UF - 1
    print('env %2d start %7d F-phase %6d gap %6s | w->conv %5d conv->mma %5d landed-w %5d | tail: wait->ready %6d | B tiles %7d..%7d drain last %7d'
          % (k, m[a], m[b] - m[a], (m[b + 1] - m[b]) if b + 1 < len(m) else '-',
             np.median(c[a:b] - w[a:b]), np.median(m[a:b] - c[a:b]), np.median(x[a:b] - w[a:b]),
             tail[2 * k + 1] - tail[2 * k] if 2 * k + 1 < len(tail) else -1,
             mb[TB * k] if TB * k < len(mb) else -1, mb[TB * k + TB - 1] if TB * k + TB - 1 < len(mb) else -1,
             dr[TB * k + TB - 1] if TB * k + TB - 1 < len(dr) else -1))
