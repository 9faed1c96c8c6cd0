"""Config 4: near-degenerate cocircular/collinear 1025x1025 lattice perturbed by 1e-15 -- every predicate falls through the
static filter; measures the exact-fallback kernels (interval, then expansions)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
import time
import numpy as np
from mmesh_b200 import capi, synth
from oracle import pyoracle as oracle
for perturb in (0.0, 1e-15):
    m = synth.config4(n=1024, perturb=perturb)
    ctx = capi.Context(2, 0)
    ctx.upload_mesh(m.xy, m.cells, m.perm, m.edges, None)
    z = np.zeros_like(m.xy)
    ctx.trial_move_validate(z); ctx.legality()
    ctx.timing_enable(True); ctx.timing_reset()
    for _ in range(5):
        v, s, ninv = ctx.trial_move_validate(z, download=False)
        f, nill = ctx.legality(download=False)
    v, s, ninv = ctx.trial_move_validate(z)
    f, nill = ctx.legality()
    res = ctx.adapt_step(z, 0)
    tm = ctx.timing_get()
    print("perturb", perturb, "cells", m.nc, "edges", m.edges.shape[0], "exact orient", res.n_exact_orient, "exact incircle", res.n_exact_incircle,
          "illegal", res.n_illegal, "nonpos", res.n_orient_nonpos)
    for name in ("validate2d", "exact_orient2d", "legality", "exact_incircle"):
        ms, cnt = tm[name]
        print("  %-16s %8.4f ms" % (name, ms / cnt))
    ctx.close()
    # the oracle's predicates (static filter, then exact expansions) over the same cells and edges, on the host
    for thr in (1, os.cpu_count()):
        oracle.set_threads(thr)
        t0 = time.perf_counter(); ov, osg, oninv = oracle.trial_validate(m.xy, None, m.cells); t1 = time.perf_counter()
        ofl, onill = oracle.legality(m.xy, None, m.edges); t2 = time.perf_counter()
        print("  oracle, %2d thread(s): orientation %.1f ms, in-circle %.1f ms -> %.3g cells/s" % (thr, 1e3 * (t1 - t0), 1e3 * (t2 - t1), m.nc / (t2 - t0)))
    oracle.set_threads(1)
    assert np.array_equal(osg, s) and np.array_equal(ofl, f)
