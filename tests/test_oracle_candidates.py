"""adapt() candidate loops of the oracle: the coarsening sort pinned against the real std::sort of libstdc++
(tests/native/sort_pin.cc), and the insert_/remove_ lists against a direct Python emulation with real sets."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "native", "_build", "sort_pin")
EV = [(0, 1), (0, 2), (1, 2)]


@pytest.fixture(scope="module")
def sort_pin():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "native", "sort_pin.cc"), "-o", EXE])

    def run(n, cases):
        txt = "%d\n" % n + "\n".join(" ".join("%d %r %d" % (l, float(x), c) for l, x, c in case) for case in cases) + "\n"
        out = subprocess.run([EXE], input=txt, capture_output=True, text=True, check=True).stdout
        return [[int(t) for t in line.split()] for line in out.strip().splitlines()]

    return run


@pytest.mark.parametrize("n", [3, 4])
def test_coarsening_sort_matches_libstdcxx(oracle, sort_pin, n):
    rng = np.random.default_rng(7 + n)
    cases = []
    for _ in range(4000):
        # few distinct values => many ties, near-ties within 1e-14 and comparator cycles
        lev = rng.integers(0, 3, n)
        ln = rng.integers(1, 4, n) * 0.25 + rng.choice([0.0, 4e-15, 2e-14, -2e-14], n)
        con = rng.integers(0, 2, n)
        cases.append(list(zip(lev.tolist(), ln.tolist(), con.tolist())))
    ref = sort_pin(n, cases)
    for case, want in zip(cases, ref):
        lev, ln, con = zip(*case)
        assert oracle.coarsening_order(lev, ln, con) == want, case


def _emulate(m, marks, le, imark, level):
    """mmesh.hh:826-908 with Python sets; coarsening choice through the pinned oracle comparator."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    v_if = np.zeros(m.nv, bool); cnt = np.zeros(m.nv, int)
    for a, b in m.segs:
        v_if[a] = v_if[b] = True; cnt[a] += 1; cnt[b] += 1
    v_bnd = np.zeros(m.nv, bool)
    for c in range(m.nc):
        for k in range(3):
            if m.nb[c][k] < 0:
                v_bnd[m.cells[c][(k + 1) % 3]] = True; v_bnd[m.cells[c][(k + 2) % 3]] = True
    iface = {(min(a, b), max(a, b)) for a, b in m.segs.tolist()}
    inserted, removed = set(), set()
    ins, remc, remv, inss, rems, remsv = [], [], [], [], [], []
    for c in range(m.nc):
        k = [int(m.cells[c][(int(m.perm[c]) >> (2 * t)) & 3]) for t in range(3)]
        if marks[c] == 1:
            a, b = k[EV[le[c]][0]], k[EV[le[c]][1]]
            key = (min(a, b), max(a, b))
            if key not in iface and key not in inserted:
                inserted.add(key); ins.append(c)
        elif marks[c] == -1:
            length = [0.0, 0.0, 0.0]
            for i, j in EV:
                d = m.xy[k[j]] - m.xy[k[i]]
                s = 0.0; s += d[0] * d[0]; s += d[1] * d[1]
                ln = float(np.sqrt(s))
                length[i] += ln; length[j] += ln
            order = pyoracle.coarsening_order([int(level[x]) for x in k], length, [bool(v_if[x] or v_bnd[x]) for x in k])
            v = k[order[0]]
            if v not in removed:
                removed.add(v); remc.append(c); remv.append(v)
    for s, (a, b) in enumerate(m.segs.tolist()):
        if imark[s] == 1:
            inss.append(s)
        elif imark[s] == -1:
            v = a if cnt[a] == 2 else (b if cnt[b] == 2 else -1)
            if v >= 0 and v not in removed:
                removed.add(v); rems.append(s); remsv.append(v)
    return dict(insert_cells=ins, remove_cells=remc, remove_vertices=remv, insert_segs=inss, remove_segs=rems,
                remove_seg_vertices=remsv)


def test_adapt_candidates_against_set_emulation(oracle):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
    from mmesh_b200 import synth
    m = synth.config1(n=24)
    rng = np.random.default_rng(3)
    dist = oracle.distance_2d(m.xy, m.segs)
    prm = oracle.indicator_init_2d(m.xy, m.edges, m.segs, oracle.Params.default())
    prm.maxH *= 0.4; prm.minH *= 1.6                      # both refine and coarsen marks present
    marks, le, counts = oracle.mark(m.xy, dist, m.cells, m.perm, prm, oracle.distance_max(dist), 2)
    assert counts[0] > 0 and counts[1] > 0
    imark = rng.integers(-1, 2, m.segs.shape[0]).astype(np.int8)
    level = rng.integers(0, 3, m.nv).astype(np.int32)
    got = oracle.adapt_candidates_2d(m.xy, m.cells, m.perm, m.nb, marks, le, m.segs, imark, level)
    want = _emulate(m, marks, le, imark, level)
    for key in want:
        assert got[key].tolist() == want[key], key
    assert len(want["remove_cells"]) > 0 and len(want["insert_cells"]) > 0


def test_refine_candidates_3d_against_set_emulation(oracle):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "dune-mmesh_b200", "python"))
    from mmesh_b200 import synth
    m = synth.lattice3d(8)
    rd = oracle.distance_3d(m.xyz, m.tris)
    prm = oracle.Params.default()
    prm.minH, prm.maxH, prm.factor = 0.25 * m.h, 1.3 * m.h, 1.5
    marks, le, counts = oracle.mark(m.xyz, rd, m.cells, m.perm, prm, oracle.distance_max(rd), dim=3)
    EV3 = [(0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3)]
    iface = set()
    for t in m.tris.tolist():
        for i, j in ((0, 1), (0, 2), (1, 2)):
            iface.add((min(t[i], t[j]), max(t[i], t[j])))
    inserted, want, want_ev = set(), [], []
    for c in range(m.nc):
        if marks[c] != 1:
            continue
        k = [int(m.cells[c][(int(m.perm[c]) >> (2 * t)) & 3]) for t in range(4)]
        a, b = k[EV3[le[c]][0]], k[EV3[le[c]][1]]
        key = (min(a, b), max(a, b))
        if key in iface or key in inserted:
            continue
        inserted.add(key); want.append(c); want_ev.append([a, b])
    got, ev = oracle.refine_candidates_3d(m.cells, m.perm, marks, le, m.tris)
    assert got.tolist() == want and ev.tolist() == want_ev
    assert 0 < len(want) < counts[0]
