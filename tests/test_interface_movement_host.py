"""Host-side branch of ensureInterfaceMovement for cells that invert under the interface movement (mmesh.hh:935-1078;
dune-mmesh_b200/host/mmesh_b200.hh::resolveInterfaceMovement2d) against the ORACLE's C restatement of the reference loop
(oracle/mmesh_oracle.c: orc_ensure_interface_movement_2d -- all elements in order, validity re-evaluated after every
move-back); a second, independent Python restatement in this file must agree with the oracle too.  CPU only: the
inverted-cell list the device would report is computed here with the same inexact area formula (the GPU run of the
whole member is tests/test_gpu_host_api.py::test_ensure_interface_movement_on_device)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "native", "_build", "interface_move_test")
EV = [(0, 1), (0, 2), (1, 2)]


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "dune-mmesh_b200", "host"),
                           os.path.join(ROOT, "tests", "native", "interface_move_test.cc"), "-o", EXE,
                           "-L", os.path.join(ROOT, "dune-mmesh_b200", "lib"), "-lmmesh_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "dune-mmesh_b200", "lib")])
    return EXE


def _perm(cell):
    order = np.argsort(cell, kind="stable")                   # TDS positions of the id-sorted corners
    return int(order[0]) | (int(order[1]) << 2) | (int(order[2]) << 4)


def _area(p, q, r):
    return ((q[0] - p[0]) * (r[1] - p[1]) - (r[0] - p[0]) * (q[1] - p[1])) / 2


def _reference_loop(x, cells, facets, vflags, levels, iverts, ishifts):
    """mmesh.hh:935-1078 restated: returns ('ok', change, removed, inserted) or ('error', message)."""
    xt = x.copy()
    idx = {int(v): i for i, v in enumerate(iverts)}
    sh = ishifts.copy()
    for v, i in idx.items():
        xt[v] = x[v] + sh[i]
    incident = {v: [] for v in range(len(x))}
    for f, (a, b) in enumerate(facets):
        incident[a].append(f); incident[b].append(f)
    segset = {(min(a, b), max(a, b)): f for f, (a, b) in enumerate(facets)}
    removed, inserted, change = [], [], False
    corner1 = lambda c0, c1: c0 + 1.0 * (c1 - c0)
    for c, cell in enumerate(cells):
        if not _area(xt[cell[0]], xt[cell[1]], xt[cell[2]]) <= 0.0:
            continue
        k = sorted(int(v) for v in cell)
        all_if = True
        for v in k:
            if not vflags[v] & 1:
                if vflags[v] & 2:
                    continue
                if v not in removed:
                    removed.append(v); change = True
                all_if = False
        if not all_if:
            continue
        ie = None
        for a, b in EV:
            if all(vflags[v] & 1 for v in (k[a], k[b])) and (min(k[a], k[b]), max(k[a], k[b])) in segset:
                if ie is not None:
                    return ("error", "more than one interface edge")
                ie = (a, b)
        if ie is None:
            non_removable = True
            for v in k:
                if len(incident[v]) == 2:
                    non_removable = False
                    if v not in removed:
                        removed.append(v); change = True
            if non_removable:
                return ("error", "without any interface edge")
            continue
        ea, eb = k[ie[0]], k[ie[1]]
        third = k[3 - ie[0] - ie[1]]
        if len(incident[third]) > 1:
            return ("error", "two interfaces cross")
        crossing = incident[third][0]
        c0 = xt[ea]; c1 = corner1(c0, xt[eb])
        ce0 = xt[facets[crossing][0]]; ce1 = corner1(ce0, xt[facets[crossing][1]])
        dP, dQ = c0 - c1, ce0 - ce1
        scale = dP[0] * dQ[1] - dQ[0] * dP[1]
        dP = dP * (ce0[0] * ce1[1] - ce0[1] * ce1[0])
        dQ = dQ * (c0[0] * c1[1] - c0[1] * c1[0])
        xi = (dQ - dP) / scale
        if (c0[0] - xi[0]) * (c1[0] - xi[0]) + (c0[1] - xi[1]) * (c1[1] - xi[1]) > 0.0:
            continue
        ln = lambda a, b: float(np.sqrt((b[0] - a[0]) ** 2 + (b[1] - a[1]) ** 2))
        inserted.append((ea, eb, float(xi[0]), float(xi[1]), third, int(levels[third]) + 1, crossing, int(ln(c0, c1) > ln(ce0, ce1))))
        change = True
        xt[third] = xt[third] - sh[idx[third]]
        sh[idx[third]] = 0.0
    return ("ok", change, removed, inserted)


def _run(exe, tmp_path, x, cells, facets, vflags, levels, iverts, ishifts):
    x = np.asarray(x, float); cells = np.asarray(cells, np.int32); facets = np.asarray(facets, np.int32).reshape(-1, 2)
    xt = x.copy()
    for i, v in enumerate(iverts):
        xt[v] = x[v] + ishifts[i]
    invalid = [c for c, cell in enumerate(cells) if _area(xt[cell[0]], xt[cell[1]], xt[cell[2]]) <= 0.0]   # what the device reports
    f = tmp_path / "snap.txt"
    with open(f, "w") as out:
        out.write("%d %d %d %d %d\n" % (len(x), len(cells), len(facets), len(iverts), len(invalid)))
        for row in x: out.write("%.17g %.17g\n" % tuple(row))
        for row in cells: out.write("%d %d %d\n" % tuple(row))
        out.write(" ".join(str(_perm(c)) for c in cells) + "\n")
        for row in facets: out.write("%d %d\n" % tuple(row))
        out.write(" ".join(str(int(v)) for v in vflags) + "\n")
        out.write(" ".join(str(int(v)) for v in levels) + "\n")
        out.write(" ".join(str(int(v)) for v in iverts) + "\n")
        for row in ishifts: out.write("%.17g %.17g\n" % tuple(row))
        out.write(" ".join(str(c) for c in invalid) + "\n")
    res = subprocess.run([exe, str(f)], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stderr
    twin = _reference_loop(x, cells, [tuple(r) for r in facets.tolist()], list(vflags), list(levels), list(iverts), np.asarray(ishifts, float))
    from oracle import pyoracle as O
    O.build()
    want = O.ensure_interface_movement_2d(x, cells, [_perm(c) for c in cells], vflags, facets, iverts, ishifts, levels)
    if want[0] == "error":
        want = ("error", {-2: "more than one interface edge", -3: "without any interface edge", -4: "two interfaces cross"}[want[1]])
    assert want == twin, (want, twin)
    return res.stdout.strip().splitlines(), want, invalid


# A(0,0) C(.5,-1) B(1,0) S(.5,1) T(.5,.3): cells ACB, ABT, BST, SAT (all ccw)
X = [(0.0, 0.0), (0.5, -1.0), (1.0, 0.0), (0.5, 1.0), (0.5, 0.3)]
CELLS = [(0, 1, 2), (0, 2, 4), (2, 3, 4), (3, 0, 4)]


def test_tip_crossing_an_interface_edge_inserts_the_intersection(exe, tmp_path):
    facets = [(0, 2), (3, 4)]
    vflags = [1, 0, 1, 1, 1]
    levels = [0, 0, 0, 1, 2]
    iverts = [0, 2, 3, 4]
    ish = np.zeros((4, 2)); ish[3] = (0.0, -0.5)                 # the tip T moves through the edge AB
    lines, want, invalid = _run(exe, tmp_path, X, CELLS, facets, vflags, levels, iverts, ish)
    assert invalid == [1] and want[0] == "ok"
    assert lines[0] == "change 1" and lines[1].split()[1:] == [] and lines[2] == "inserted 1"
    ip = lines[3].split()
    assert [int(ip[1]), int(ip[2])] == [0, 2] and int(ip[5]) == 4 and int(ip[6]) == 3 and int(ip[7]) == 1 and int(ip[8]) == 0
    assert (float(ip[3]), float(ip[4])) == (want[3][0][2], want[3][0][3]) == (0.5, 0.0)


def test_non_interface_corners_of_inverted_cells_are_removed_once(exe, tmp_path):
    facets = [(0, 2)]
    vflags = [1, 0, 1, 0, 0]
    iverts = [0, 2]
    ish = np.array([(0.6, 0.4), (0.0, 0.0)])                     # A moves into the fan around T
    lines, want, invalid = _run(exe, tmp_path, X, CELLS, facets, vflags, [0] * 5, iverts, ish)
    assert want[0] == "ok" and len(invalid) >= 2
    assert lines[0] == "change %d" % int(want[1])
    assert [int(v) for v in lines[1].split()[1:]] == want[2] == [4, 3]
    assert lines[2] == "inserted 0"


def test_boundary_flagged_corner_is_kept(exe, tmp_path):
    facets = [(0, 2)]
    vflags = [1, 0, 1, 2, 0]                                     # S has boundaryFlag == 1: not removable
    lines, want, _ = _run(exe, tmp_path, X, CELLS, facets, vflags, [0] * 5, [0, 2], np.array([(0.6, 0.4), (0.0, 0.0)]))
    assert [int(v) for v in lines[1].split()[1:]] == want[2] == [4]


@pytest.mark.parametrize("facets,vflags,iverts,message", [
    ([(0, 2), (2, 4)], [1, 0, 1, 0, 1], [0, 2, 4], "more than one interface edge"),
    ([(0, 1), (1, 2), (3, 4)], [1, 1, 1, 1, 1], [0, 1, 2, 3, 4], "without any interface edge"),
    ([(0, 2), (3, 4), (4, 2)], [1, 0, 1, 1, 1], [0, 2, 3, 4], "more than one interface edge"),
])
def test_unresolvable_constrained_cells_raise_grid_error(exe, tmp_path, facets, vflags, iverts, message):
    ish = np.zeros((len(iverts), 2)); ish[iverts.index(4)] = (0.0, -0.5)
    lines, want, invalid = _run(exe, tmp_path, X, CELLS, facets, vflags, [0] * 5, iverts, ish)
    assert invalid == [1]
    assert want == ("error", message)
    assert lines[0].startswith("GridError") and message in lines[0]


def test_random_movements_match_the_reference_loop(exe, tmp_path):
    """a strip of cells above an interface line with several interface tips hanging into it; random tip movements"""
    rng = np.random.default_rng(4)
    n = 8
    # bottom row y=0 (interface line), top row y=1, tips at y=0.3 between them
    bottom = [(float(i), 0.0) for i in range(n + 1)]
    top = [(float(i), 1.0) for i in range(n + 1)]
    tips = [(i + 0.5, 0.3) for i in range(n)]
    x = np.array(bottom + top + tips)
    B = lambda i: i
    Tp = lambda i: n + 1 + i
    Ti = lambda i: 2 * (n + 1) + i
    cells, facets = [], []
    for i in range(n):
        cells += [(B(i), B(i + 1), Ti(i)), (B(i + 1), Tp(i + 1), Ti(i)), (Tp(i + 1), Tp(i), Ti(i)), (Tp(i), B(i), Ti(i))]
        facets.append((B(i), B(i + 1)))
        facets.append((Ti(i), Tp(i)) if i % 2 == 0 else (Tp(i + 1), Ti(i)))      # a second interface ending in the tip
    vflags = np.zeros(len(x), int)
    for a, b in facets:
        vflags[a] |= 1; vflags[b] |= 1
    iverts = [int(v) for v in np.nonzero(vflags & 1)[0]]
    levels = rng.integers(0, 3, len(x))
    checked = 0
    for trial in range(40):
        ish = np.zeros((len(iverts), 2))
        for i in range(n):
            if rng.uniform() < 0.5:
                ish[iverts.index(Ti(i))] = (rng.uniform(-0.3, 0.3), -rng.uniform(0.2, 0.8))
        lines, want, invalid = _run(exe, tmp_path, x, cells, facets, vflags, levels, iverts, ish)
        if want[0] == "error":
            assert lines[0].startswith("GridError") and want[1] in lines[0]
            continue
        assert lines[0] == "change %d" % int(want[1])
        assert [int(v) for v in lines[1].split()[1:]] == want[2]
        assert lines[2] == "inserted %d" % len(want[3])
        for ln, w in zip(lines[3:], want[3]):
            t = ln.split()
            assert [int(t[1]), int(t[2])] == [w[0], w[1]] and (float(t[3]), float(t[4])) == (w[2], w[3])
            assert [int(t[5]), int(t[6]), int(t[7]), int(t[8])] == [w[4], w[5], w[6], w[7]]
        checked += len(want[3])
    assert checked > 10
