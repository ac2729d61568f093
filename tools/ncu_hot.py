#!/usr/bin/env python
"""Top stall sites of a kernel from `ncu -i X.ncu-rep --page source --csv` (needs -lineinfo / --import-source).

    python tools/ncu_hot.py gpurun_out/prof_X.ncu-rep [N]
"""
import csv
import subprocess
import sys

raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == 'Kernel Name':
        name = rows[k][1]
        h = rows[k + 1]
        ia, isrc, ismp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
        stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
        data, tot, base = [], 0, None
        k += 2
        while k < len(rows) and not (rows[k] and rows[k][0] == 'Kernel Name'):
            r = rows[k]
            k += 1
            try:
                n = int(r[ismp])
                a = int(r[ia], 16)
            except (ValueError, IndexError):
                continue
            base = a if base is None else base
            tot += n
            st = ' '.join('%s=%s' % (h[i][6:], r[i]) for i in stall if r[i] not in ('0', ''))
            data.append((n, a - base, r[isrc].strip(), st, r[iex]))
        print('## %s: %d samples' % (name, tot))
        agg = {}
        for n, a, s, st, ex in data:
            for kv in st.split():
                kk, vv = kv.split('=')
                agg[kk] = agg.get(kk, 0) + int(vv)
        print('by reason:', ' '.join('%s=%d' % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])))
        for n, a, s, st, ex in sorted(data, key=lambda d: -d[0])[:top]:
            print('%4d 0x%04x %-64s ex=%-7s %s' % (n, a, s[:64], ex, st))
        break
    k += 1
