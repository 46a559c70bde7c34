#!/usr/bin/env python
"""Per-kernel hot SASS lines from an .ncu-rep holding several kernels: ncu_lines.py rep kernel_substr [min_pct]"""
import csv, subprocess, sys
rep, sub = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; blocks.append(cur)
    elif cur is not None:
        cur['rows'].append(r)
for b in blocks:
    if sub not in b['name']:
        continue
    hdr = b['rows'][0]; ix = {h: i for i, h in enumerate(hdr)}; body = b['rows'][1:]
    tot = sum(int(r[ix['# Samples']] or 0) for r in body)
    print("==", b['name'][:70], "samples", tot, "instr", sum(int(r[ix['Instructions Executed']] or 0) for r in body))
    sc = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    for pos, r in enumerate(body):
        s = int(r[ix['# Samples']] or 0)
        if s > tot * minpct / 100:
            st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in sc), reverse=True)[:2]
            print("%5d %5.1f%% exe=%9s thr=%3s %-58s %s" % (pos, 100 * s / tot, r[ix['Instructions Executed']], r[ix['Avg. Threads Executed']], r[ix['Source']].strip()[:58], st))
    break
