"""adapt_()'s component numbering (mmesh.hh:1194-1218 with getComponentNumber_ :1919-1941): the oracle restates the
order-dependent `while (markAgain)` loop statement by statement; the closed form the device evaluates (fresh zones numbered
in order, group minimum by label propagation) must agree with it on every input."""
import numpy as np


def random_zones(rng, nc, nz, max_size=6):
    sizes = rng.integers(1, max_size + 1, nz)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    cells = np.concatenate([rng.choice(nc, s, replace=False) for s in sizes]).astype(np.int32) if nz else np.zeros(0, np.int32)
    return off, cells


def closed_form(zone_off, zone_cells, nc, base=1):
    nz = len(zone_off) - 1
    zid = np.repeat(np.arange(nz), np.diff(zone_off))
    first = np.full(nc, nz, np.int64)
    np.minimum.at(first, zone_cells, zid)
    fresh = np.ones(nz, bool)
    np.logical_and.at(fresh, zid, first[zone_cells] == zid)
    big = 1 << 40
    label = np.where(fresh, base + np.cumsum(fresh) - 1, big)
    cl = np.full(nc, big, np.int64)
    while True:
        np.minimum.at(cl, zone_cells, label[zid])
        new = label.copy()
        np.minimum.at(new, zid, cl[zone_cells])
        if np.array_equal(new, label):
            break
        label = new
    cl[cl == big] = 0
    return label.astype(np.int32), cl.astype(np.int32), base + int(fresh.sum())


def test_reference_loop_examples(oracle):
    # zones: {0,1} {2,3} {1,2} -> the third zone merges the first two into number 1; a later disjoint zone gets 3
    off = np.array([0, 2, 4, 6, 7], np.int64)
    cells = np.array([0, 1, 2, 3, 1, 2, 5], np.int32)
    zn, cn, cc = oracle.component_numbers(off, cells, 6, 1)
    assert zn.tolist() == [1, 1, 1, 3] and cn.tolist() == [1, 1, 1, 1, 0, 3] and cc == 4
    # componentCount_ continues from connectedComponents_.size() + 1
    zn, cn, cc = oracle.component_numbers(off, cells, 6, 10)
    assert zn.tolist() == [10, 10, 10, 12] and cc == 13


def test_closed_form_equals_reference_loop(oracle):
    rng = np.random.default_rng(0)
    for trial in range(400):
        nc = int(rng.integers(3, 80)); nz = int(rng.integers(0, 40))
        off, cells = random_zones(rng, nc, nz, max_size=min(6, nc))
        base = int(rng.integers(1, 5))
        zn, cn, cc = oracle.component_numbers(off, cells, nc, base)
        l, c, k = closed_form(off, cells, nc, base)
        assert np.array_equal(zn, l) and np.array_equal(cn, c) and cc == k
