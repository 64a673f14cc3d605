import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, sys
from oracle import oracle
from tests.emul import emul
from tests.helpers import GOLDEN_CASES, load_case
for name in GOLDEN_CASES:
    scene, rays, _, gl = load_case(name)
    want, _ = oracle.trace(scene, rays, gl)
    got = emul.trace_f32(scene, rays, gl)
    if got is None:
        print(name, 'generic: unsupported'); continue
    import collections
    bad = []
    for i in range(rays.shape[1]):
        w = want[5, want[4]==i]; g = got[5, got[4]==i]
        if w.shape != g.shape or not np.array_equal(w, g): bad.append(i)
    if bad:
        print(name, 'rows', want.shape[1], got.shape[1], 'rays with different surface sequence', len(bad), 'of', rays.shape[1], bad[:6])
    keep_w = ~np.isin(want[4], bad); keep_g = ~np.isin(got[4], bad)
    w, g = want[:, keep_w], got[:, keep_g]
    scale = max(1.0, np.nanmax(np.abs(w[6:12])))
    perr = np.nanmax(np.abs(g[6:12]-w[6:12]))/scale
    derr = np.nanmax(np.abs(g[12:15]-w[12:15]))
    ierr = np.nanmax(np.abs(g[3]-w[3]))
    print(name, 'rows', w.shape[1], 'pos err/scale %.2e dir err %.2e idx err %.2e scale %.0f' % (perr, derr, ierr, scale))
