import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle import oracle as O
import parity as P
d = O.read_case('/root/repo/tests/golden/ermak')
o = O.Oracle(**d)
ls = P.Lockstep(o, strict=1)
for i in range(1000):
    try:
        ls.step(check=True, tag="step %d" % (i + 1))
    except AssertionError as e:
        print("FAIL", str(e)[:600])
        a = P.oracle_slot_arrays(o); g = ls.ctx.download(len(a['z']))
        for k in ('pos','vel','acel','pos_old','old_cg','z','flags'):
            x, y = g[k], a[k]
            bad = np.flatnonzero((x != y).reshape(len(x), -1).any(axis=1) & a['alive'])
            print(k, bad[:10])
            for b in bad[:4]:
                print('   slot', b, 'gpu', x[b], 'orc', y[b], 'z', a['z'][b], 'flags', a['flags'][b], g['flags'][b])
        s = o.scalars(); c = ls.ctx.counters()
        print('choques', s.choques, c.choques, 'choques3', s.choques3, c.choques3, 'zmax', s.zmax, ls.ctx.scalars().zmax)
        break
else:
    print("no failure in 1000 steps")
