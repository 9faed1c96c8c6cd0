"""The C++ host class (Dune::B200::MMesh, reference member names) driven like the reference's user loop, checked
against the oracle.  Compiled with g++ against the C-ABI library."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "native", "_build", "host_api_test")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "dune-mmesh_b200", "host"),
                           os.path.join(ROOT, "tests", "native", "host_api_test.cc"), "-o", EXE,
                           "-L", os.path.join(ROOT, "dune-mmesh_b200", "lib"), "-lmmesh_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "dune-mmesh_b200", "lib")])


def test_host_class_compiles_and_links():
    from mmesh_b200 import capi
    capi.build()
    build_exe()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_class_user_loop(oracle, tmp_path):
    from mmesh_b200 import synth
    build_exe()
    m = synth.config1(n=24)
    s = synth.shifts_uniform(m, 0.9 * m.h, 31)
    f = tmp_path / "snap.txt"
    with open(f, "w") as out:
        out.write("%d %d %d %d\n" % (m.nv, m.nc, m.edges.shape[0], m.segs.shape[0]))
        for arr, fmt in ((m.xy, "%.17g"), (s, "%.17g"), (m.vflags, "%d"), (m.cells, "%d"), (m.perm, "%d"), (m.edges, "%d"), (m.segs, "%d"), (m.nb, "%d")):
            np.savetxt(out, np.asarray(arr).reshape(len(arr), -1), fmt=fmt)
    res = subprocess.run([EXE, str(f)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    lines = {l.split()[0]: l.split()[1:] for l in res.stdout.strip().splitlines()}
    prm = oracle.indicator_init_2d(m.xy, m.edges, m.segs, oracle.Params.default())
    assert [float(v) for v in lines["params"]] == [prm.minH, prm.maxH, prm.factor]
    prm.maxH *= 0.45
    prm.minH *= 3.0
    d = oracle.distance_2d(m.xy, m.segs)
    marks, le, counts = oracle.mark(m.xy, d, m.cells, m.perm, prm, oracle.distance_max(d))
    assert np.array_equal(np.array(lines["marks"], int), marks)
    assert np.array_equal(np.array(lines["longest"], int), le)
    assert float(lines["maxdist"][0]) == oracle.distance_max(d)
    assert int(lines["marked"][0]) == int(counts[0] + counts[1] > 0 or True)
    imark = oracle.mark_interface_2d(m.xy, m.segs, prm, oracle.distance_max(d))
    cand = oracle.adapt_candidates_2d(m.xy, m.cells, m.perm, m.nb, marks, le, m.segs, imark, None)
    assert [int(v) for v in lines["insertCells"]] == cand["insert_cells"].tolist()
    assert [int(v) for v in lines["removeCells"]] == cand["remove_cells"].tolist()
    assert [int(v) for v in lines["removeVertices"]] == cand["remove_vertices"].tolist()
    assert [int(v) for v in lines["insertSegments"]] == cand["insert_segs"].tolist()
    assert [int(v) for v in lines["removeSegmentVertices"]] == cand["remove_seg_vertices"].tolist()
    EV = [(0, 1), (0, 2), (1, 2)]
    pts = np.array([float(v) for v in lines["insertPoints"]]).reshape(-1, 2)
    for p, c in zip(pts, cand["insert_cells"]):
        k = [int(m.cells[c][(int(m.perm[c]) >> (2 * t)) & 3]) for t in range(3)]
        a, b = m.xy[k[EV[le[c]][0]]], m.xy[k[EV[le[c]][1]]]
        assert np.array_equal(p, a + 0.5 * (b - a))
    assert len(cand["insert_cells"]) + len(cand["remove_cells"]) > 0
    valid, _, ninv = oracle.trial_validate(m.xy, s, m.cells)
    # host part of ensureVertexMovement (mmesh.hh:1109-1124): non-interface, non-boundary corners of invalid cells,
    # in sub-entity (id) order, first occurrence only
    removed = []
    for c in np.nonzero(valid == 0)[0]:
        for k in range(3):
            v = int(m.cells[c][(m.perm[c] >> (2 * k)) & 3])
            if m.vflags[v] & 3:
                continue
            if v not in removed:
                removed.append(v)
    assert [int(v) for v in lines.get("removed", [])] == removed and int(lines["change"][0]) == int(len(removed) > 0)
    assert ninv > 0
    legal, _ = oracle.legality(oracle.move(m.xy, s), None, m.edges)
    assert np.array_equal(np.array(lines["legal"], int), legal)
    assert lines["preAdapt"] == ["0"] and lines["sizecheck"] == ["ok"]
    # the fused step() call returns the same fields as the member-by-member loop
    assert np.array_equal(np.array(lines["step_marks"], int), marks)
    assert np.array_equal(np.array(lines["step_valid"], int), valid)
    # move_policy 0 (default): the trial move inverts cells => moveVertices withheld on the device, legality describes the
    # mesh as it stays; the next step with move_policy 1 applies the same shifts
    legal_unmoved, _ = oracle.legality(m.xy, None, m.edges)
    assert lines["step_moved"] == ["0"]
    assert np.array_equal(np.array(lines["step_legal"], int), legal_unmoved)
    assert [int(v) for v in lines["step_counts"]] == [int(counts[0]), int(counts[1]), int(ninv), int((legal_unmoved == 0).sum())]
    assert lines["step1_moved"] == ["1", str(int(ninv))]
    assert np.array_equal(np.array(lines["step1_legal"], int), legal)
    # after the (large) move the mesh has inverted cells => ensureInterfaceMovement's pre-check throws (mmesh.hh:926-930)
    _, _, ninv_moved = oracle.trial_validate(oracle.move(m.xy, s), None, m.cells)
    assert lines["ifmove"][0] == ("GridError" if ninv_moved > 0 else "ok")
    # restoreDelaunay on the unmoved snapshot: the snapshot object carries the flipped topology
    oc, onb, oflips = oracle.restore_delaunay_2d(m.xy, m.cells, m.nb, m.segs)
    fl = [int(v) for v in lines["flip"]]
    fc = np.array(lines["flipcells"], int).reshape(-1, 3)
    tri = lambda c: set(map(tuple, np.sort(c, axis=1).tolist()))
    assert fl[0] > 0 and oflips > 0 and fl[2] == 1 and tri(fc) == tri(oc)
    fe = np.array(lines["flipedges"], int).reshape(-1, 4)
    assert np.array_equal(fe, synth.build_edges(fc.astype(np.int32), m.nv)[0])
    oflags, onill = oracle.legality(m.xy, None, fe.astype(np.int32))
    assert int(lines["flipillegal"][0]) == onill == fl[3]
    assert lines["flipinverted"] == ["GridError"]
    assert [int(v) for v in lines["fliplog"]] == [fl[0], fl[0], fl[1] + 1]            # total, logged entries, rounds + 1 offsets


ADAPT_EXE = os.path.join(ROOT, "tests", "native", "_build", "adapt_handle_test")


def build_adapt_exe():
    os.makedirs(os.path.dirname(ADAPT_EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "dune-mmesh_b200", "host"),
                           os.path.join(ROOT, "tests", "native", "adapt_handle_test.cc"), "-o", ADAPT_EXE,
                           "-L", os.path.join(ROOT, "dune-mmesh_b200", "lib"), "-lmmesh_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "dune-mmesh_b200", "lib")])


def test_adapt_handle_driver_compiles_and_links():
    from mmesh_b200 import capi
    capi.build()
    build_adapt_exe()
    assert os.path.exists(ADAPT_EXE)


@pytest.mark.gpu
@pytest.mark.parametrize("ncomp", [1, 3])
def test_adapt_handle_drop_in(oracle, tmp_path, ncomp):
    """MMesh::adapt(handle) (mmesh.hh:1138-1165) through the C++ host class: old snapshot -> TDS edit (generated
    insertions / removals) -> new snapshot; pairs from the two snapshots, gather -> device projection -> scatter.
    Expected values: the oracle's transfer over the pair list of a brute-force overlap search (every vanished cell that
    overlaps a new cell), which is independent of the product's pair builder."""
    from mmesh_b200 import synth
    from adapt_cases import adapt_case
    build_adapt_exe()
    case = adapt_case(n=16, seed=21, n_insert=30, n_remove=30)
    old = case["old"]
    k = np.stack([np.take_along_axis(old.cells, ((old.perm >> (2 * t)) & 3)[:, None].astype(np.int64), 1)[:, 0] for t in range(3)], 1)
    pk = old.xy[k]
    if ncomp == 3:
        dofs_old = np.sin(np.pi * pk[..., 0]) * np.sin(np.pi * pk[..., 1]) + 2.0
    else:
        dofs_old = (np.sin(np.pi * pk[..., 0]) * np.sin(np.pi * pk[..., 1]) + 2.0).mean(axis=1, keepdims=True)
    ncn = case["new_cells"].shape[0]
    carried = np.full((ncn, ncomp), -7.0)                      # unchanged cells: whatever the handle holds stays
    new_edges, _, _ = synth.build_edges(case["new_cells"], case["new_xy"].shape[0])
    f = tmp_path / "adapt.bin"
    with open(f, "wb") as out:
        for (xy, ids, cells, perm, edges) in ((old.xy, case["old_id"], old.cells, old.perm, old.edges),
                                              (case["new_xy"], case["new_id"], case["new_cells"], case["new_perm"], new_edges)):
            np.array([xy.shape[0], cells.shape[0], edges.shape[0], 0], np.int64).tofile(out)
            np.ascontiguousarray(xy, np.float64).tofile(out); np.ascontiguousarray(ids, np.int64).tofile(out)
            np.ascontiguousarray(cells, np.int32).tofile(out); np.ascontiguousarray(perm, np.uint8).tofile(out)
            np.ascontiguousarray(edges, np.int32).tofile(out)
        np.array([ncomp], np.int64).tofile(out)
        np.ascontiguousarray(dofs_old, np.float64).tofile(out); carried.tofile(out)
    res = subprocess.run([ADAPT_EXE, str(f)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    lines = {l.split()[0]: l.split()[1:] for l in res.stdout.strip().splitlines()}
    got = np.array([float(v) for v in lines["dofs"]]).reshape(ncn, ncomp)
    created = np.nonzero(case["created"])[0]
    assert int(lines["adapt"][0]) == 1 and int(lines["adapt"][2]) == 1 and int(lines["adapt"][6]) == created.size
    assert np.array_equal(got[~case["created"]], carried[~case["created"]])
    # independent pair list: brute-force overlap search, children in ascending old index
    van = np.nonzero(case["vanished"])[0]
    newc = case["new_xy"][case["new_cells"]]
    off, oidx = [0], []
    for K in created:
        for o in van:
            if oracle.intersection_volume(pk[o].reshape(6), newc[K].reshape(6)) > 0.0:
                oidx.append(o)
        off.append(len(oidx))
    oidx = np.array(oidx, np.int32)
    prm = oracle.Params.default()
    ref, mass, cov, npc = oracle.project(case["new_xy"], case["new_cells"], case["new_perm"], np.array(off, np.int64), created.astype(np.int32),
                                         pk[oidx].reshape(-1, 6), oidx, dofs_old, ncn, ncomp, prm)
    ref = ref.reshape(ncn, ncomp)
    # accumulation order differs (component order vs ascending index), pieces below the 1e-8 cut-off differ by touching pairs: 1e-12
    assert np.abs(got[created] - ref[created]).max() <= 1e-12 * np.abs(ref[created]).max()
    assert int(lines["after"][1]) == 0


IFMOVE_EXE = os.path.join(ROOT, "tests", "native", "_build", "interface_move_gpu_test")


@pytest.mark.gpu
def test_ensure_interface_movement_on_device(oracle, tmp_path):
    """MMesh::ensureInterfaceMovement (mmesh.hh:920-1088) through the GPU path on a mesh that was moved since its snapshot:
    pre-check + inverted cells on the device, the element loop on the CURRENT coordinates.  Expected: the oracle's C
    restatement of the reference loop, run on the moved coordinates (moveVertices is one IEEE add per component)."""
    os.makedirs(os.path.dirname(IFMOVE_EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "dune-mmesh_b200", "host"),
                           os.path.join(ROOT, "tests", "native", "interface_move_gpu_test.cc"), "-o", IFMOVE_EXE,
                           "-L", os.path.join(ROOT, "dune-mmesh_b200", "lib"), "-lmmesh_b200", "-Wl,-rpath," + os.path.join(ROOT, "dune-mmesh_b200", "lib")])
    rng = np.random.default_rng(14)
    n = 24
    bottom = [(float(i), 0.0) for i in range(n + 1)]; top = [(float(i), 1.0) for i in range(n + 1)]; tips = [(i + 0.5, 0.3) for i in range(n)]
    x = np.array(bottom + top + tips)
    B = lambda i: i; Tp = lambda i: n + 1 + i; Ti = lambda i: 2 * (n + 1) + i
    cells, facets = [], []
    for i in range(n):
        cells += [(B(i), B(i + 1), Ti(i)), (B(i + 1), Tp(i + 1), Ti(i)), (Tp(i + 1), Tp(i), Ti(i)), (Tp(i), B(i), Ti(i))]
        facets.append((B(i), B(i + 1)))
        facets.append((Ti(i), Tp(i)) if i % 2 == 0 else (Tp(i + 1), Ti(i)))
    cells = np.array(cells, np.int32); facets = np.array(facets, np.int32)
    perm = np.array([int(o[0]) | (int(o[1]) << 2) | (int(o[2]) << 4) for o in np.argsort(cells, axis=1, kind="stable")], np.uint8)
    vflags = np.zeros(len(x), np.uint8)
    vflags[facets.ravel()] |= 1
    iverts = np.nonzero(vflags & 1)[0].astype(np.int32)
    levels = rng.integers(0, 3, len(x)).astype(np.int32)
    seen_insert = seen_error = 0
    for trial in range(12):
        vsh = rng.uniform(-0.05, 0.05, x.shape)                       # the mesh moves first: the snapshot coordinates go stale
        vsh[: 2 * (n + 1), 1] = 0.0
        ish = np.zeros((len(iverts), 2))
        for i in range(n):
            if rng.uniform() < 0.4:
                ish[list(iverts).index(Ti(i))] = (rng.uniform(-0.2, 0.2), -rng.uniform(0.25, 0.7))
        f = tmp_path / ("ifm_%d.txt" % trial)
        with open(f, "w") as out:
            out.write("%d %d %d %d\n" % (len(x), len(cells), len(facets), len(iverts)))
            for arr, fmt in ((x, "%.17g"), (cells, "%d")):
                np.savetxt(out, arr, fmt=fmt)
            out.write(" ".join(str(int(p)) for p in perm) + "\n")
            np.savetxt(out, facets, fmt="%d")
            for arr in (vflags, levels, iverts):
                out.write(" ".join(str(int(v)) for v in arr) + "\n")
            np.savetxt(out, ish, fmt="%.17g"); np.savetxt(out, vsh, fmt="%.17g")
        res = subprocess.run([IFMOVE_EXE, str(f)], capture_output=True, text=True, timeout=120)
        assert res.returncode == 0, res.stderr
        lines = res.stdout.strip().splitlines()
        want = oracle.ensure_interface_movement_2d(oracle.move(x, vsh), cells, perm, vflags, facets, iverts, ish, levels)
        if want[0] == "error":
            msg = {-1: "negative volume", -2: "more than one interface edge", -3: "without any interface edge", -4: "two interfaces cross"}[want[1]]
            assert lines[0].startswith("GridError") and msg in lines[0]
            seen_error += 1
            continue
        assert lines[0] == "change %d" % int(want[1])
        assert [int(v) for v in lines[1].split()[1:]] == want[2]
        assert lines[2] == "inserted %d" % len(want[3])
        for ln, w in zip(lines[3:], want[3]):
            t = ln.split()
            assert [int(t[1]), int(t[2])] == [w[0], w[1]] and (float(t[3]), float(t[4])) == (w[2], w[3])
            assert [int(t[5]), int(t[6]), int(t[7]), int(t[8])] == [w[4], w[5], w[6], w[7]]
        assert lines[3 + len(want[3])] == "coords_kept 1"
        seen_insert += len(want[3])
    assert seen_insert > 3
