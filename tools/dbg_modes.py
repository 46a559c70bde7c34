import os, sys, subprocess
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from ufomap_b200 import scans, capi

def run(mode, bricks, color=False):
    env = dict(os.environ); env['UFO_B200_MARK'] = mode
    code = r'''
import sys; sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests")
import numpy as np
from ufomap_b200 import scans, capi
color = %r
m = capi.Map(0.02 if not color else 0.01, color=color, initial_bricks=%d)
for k in range(2):
    if color:
        o,p,c = scans.rgbd(k=k, width=160, height=120); m.insert(o,p,rgb=c,max_range=5.0,discrete=True)
    else:
        o,p = scans.velodyne64(k=k, rings=16, azimuths=1024); m.insert(o,p,max_range=30.0)
    print("scan",k,"regrows",m.stats()["regrows"],"touched",m.stats()["touched_voxels"], file=sys.stderr)
c,v,g = m.value_field()
np.savez("/tmp/f_%s_%d_%d.npz", c=c, v=v, g=g)
''' % (color, bricks, mode, bricks, int(color))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
    print(mode, bricks, color, r.stderr.strip().replace('\n', ' | ')[-300:])
    return np.load("/tmp/f_%s_%d_%d.npz" % (mode, bricks, int(color)))

for color in (False, True):
    base = run('records', 1 << 18, color)
    for mode, bricks in (('probe', 1 << 18), ('dense', 1 << 18), ('probe', 512), ('dense', 512), ('records', 512)):
        f = run(mode, bricks, color)
        same = len(f['c']) == len(base['c']) and np.array_equal(f['c'], base['c']) and np.array_equal(f['v'].view(np.uint32), base['v'].view(np.uint32)) and np.array_equal(f['g'], base['g'])
        print('   ->', mode, bricks, 'color' if color else 'mono', 'SAME' if same else 'DIFF', len(f['c']), len(base['c']))
        if not same:
            only_b = np.setdiff1d(base['c'], f['c']); only_f = np.setdiff1d(f['c'], base['c'])
            print('      missing', len(only_b), 'extra', len(only_f))
            if len(only_b):
                vb = base['v'][np.searchsorted(base['c'], only_b[:2000])]
                print('      missing values', np.unique(vb)[:5], 'first codes', [hex(int(x)) for x in only_b[:4]])
            common = np.intersect1d(base['c'], f['c'])
            a = base['v'][np.searchsorted(base['c'], common)]; b = f['v'][np.searchsorted(f['c'], common)]
            print('      value diffs on common', int((a.view(np.uint32) != b.view(np.uint32)).sum()))
            ga = base['g'][np.searchsorted(base['c'], common)]; gb = f['g'][np.searchsorted(f['c'], common)]
            print('      colour diffs on common', int((ga != gb).any(axis=1).sum()))
