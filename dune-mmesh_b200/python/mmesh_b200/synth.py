"""Synthetic mesh snapshots for the benchmark / parity configurations.

These generators stand in for the reference's mesh readers (dune/mmesh/grid/gmshgridfactory.hh,
explicitgridfactory.hh) whose .msh inputs are git-LFS pointers.  They emit exactly the SoA snapshot a
host adapter would extract from a live Dune::MMesh:

  xy / xyz  vertex coordinates in leaf-vertex-index order        (indexsets.hh:242-245)
  vid       persistent vertex ids (VertexInfo::id)                 (cgal/defaults.hh:41)
  cells     face->vertex(k)->info().index, CGAL TDS (ccw) order    (CGAL/Triangulation_2.h:1152-1154)
  perm      ElementInfo::cgalIndex packed 2 bits/corner             (common.hh:55-65)
  edges     (a, b, c, d): edge (a,b), left apex c (a,b,c ccw), opposite apex d or -1
  segs      interface facets as vertex-index pairs, lower persistent id first (geometry.hh:88-104)

Randomness is splitmix64 keyed by (seed, persistent vertex id, component) so that every run, rank
and language sees identical bits.
"""
import numpy as np

_U64 = np.uint64


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        z += _U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        z = z ^ (z >> _U64(31))
    return z


def uniform01(seed, ids, comp):
    """U[0,1) keyed by (seed, id, component)."""
    ids = np.asarray(ids, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = _U64(seed) * _U64(0xD1342543DE82EF95) + ids * _U64(4) + _U64(comp)
    z = splitmix64(splitmix64(key))
    return (z >> _U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def cross2(a, b):
    return a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]


def hilbert_key(order, x, y):
    """Vectorised Hilbert index of integer points on a 2^order grid."""
    x = np.asarray(x, dtype=np.int64).copy()
    y = np.asarray(y, dtype=np.int64).copy()
    n = np.int64(1) << order
    d = np.zeros_like(x)
    s = n >> 1
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, n - 1 - x, x)
        y = np.where(flip, n - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return d


def pack_perm(ids_of_corners):
    """perm byte: bits (2k+1:2k) = TDS position of the k-th smallest persistent id (computeCGALIndices)."""
    order = np.argsort(ids_of_corners, axis=1, kind="stable").astype(np.uint8)
    p = np.zeros(order.shape[0], np.uint8)
    for k in range(order.shape[1]):
        p |= (order[:, k] << np.uint8(2 * k)).astype(np.uint8)
    return p


def build_edges(cells, nv):
    """Unique edges with both apices, ordered by (left cell, local index).  Also returns face neighbours."""
    nc = cells.shape[0]
    c = cells.astype(np.int64)
    a = np.concatenate([c[:, 1], c[:, 2], c[:, 0]])   # half-edge k: (v[k+1], v[k+2]), apex v[k]
    b = np.concatenate([c[:, 2], c[:, 0], c[:, 1]])
    apex = np.concatenate([c[:, 0], c[:, 1], c[:, 2]])
    cell_of = np.tile(np.arange(nc, dtype=np.int64), 3)
    key = np.minimum(a, b) * np.int64(nv) + np.maximum(a, b)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.ones(ks.shape[0], bool)
    first[1:] = ks[1:] != ks[:-1]
    has_twin = np.zeros(ks.shape[0], bool)
    has_twin[:-1] = ks[:-1] == ks[1:]
    idx_first = np.nonzero(first)[0]
    he0 = order[idx_first]                        # stable sort: lower half-edge id first => lower cell first per k-block
    twin_ok = has_twin[idx_first]
    he1 = np.where(twin_ok, order[np.minimum(idx_first + 1, order.shape[0] - 1)], -1)
    # choose the half-edge of the lower-index cell as "left"
    c0 = cell_of[he0]
    c1 = np.where(he1 >= 0, cell_of[np.maximum(he1, 0)], np.int64(1) << 60)
    swap = c1 < c0
    left = np.where(swap, he1, he0)
    right = np.where(swap, he0, he1)
    edges = np.empty((left.shape[0], 4), np.int32)
    edges[:, 0] = a[left]
    edges[:, 1] = b[left]
    edges[:, 2] = apex[left]
    edges[:, 3] = np.where(right >= 0, apex[np.maximum(right, 0)], -1)
    lc = cell_of[left]
    lk = left // nc
    o = np.lexsort((lk, lc))
    edges = edges[o]
    left, right = left[o], right[o]
    edge_cell = lc[o].astype(np.int32)
    nb = np.full((nc, 3), -1, np.int32)
    inner = right >= 0
    nb[cell_of[left[inner]], left[inner] // nc] = cell_of[right[inner]]
    nb[cell_of[right[inner]], right[inner] // nc] = cell_of[left[inner]]
    return edges, nb, edge_cell


class Mesh2D:
    dim = 2

    def __init__(self):
        self.xy = self.vid = self.vflags = self.cells = self.perm = self.edges = self.nb = self.segs = self.edge_cell = None
        self.h = None
        self.lat = None      # (i, j) lattice coordinates per vertex (index order), for analytic fields

    @property
    def nv(self):
        return self.xy.shape[0]

    @property
    def nc(self):
        return self.cells.shape[0]


def trace_curve(curve_xy, hx, hy, nx, ny, closed):
    """Snap a densely sampled curve to lattice vertices and return an 8-connected simple vertex path."""
    ij = np.stack([np.rint(curve_xy[:, 0] / hx), np.rint(curve_xy[:, 1] / hy)], axis=1).astype(np.int64)
    ij[:, 0] = np.clip(ij[:, 0], 0, nx)
    ij[:, 1] = np.clip(ij[:, 1], 0, ny)
    keep = np.ones(ij.shape[0], bool)
    keep[1:] = np.any(ij[1:] != ij[:-1], axis=1)
    path = [tuple(p) for p in ij[keep]]
    if closed and len(path) > 1 and path[0] == path[-1]:
        path.pop()
    # bridge gaps (should not occur with dense sampling) and remove spikes A,B,A
    out = []
    for p in path:
        while out and max(abs(p[0] - out[-1][0]), abs(p[1] - out[-1][1])) > 1:
            q = out[-1]
            out.append((q[0] + int(np.sign(p[0] - q[0])), q[1] + int(np.sign(p[1] - q[1]))))
        if len(out) >= 2 and out[-2] == p:
            out.pop()
            continue
        if out and out[-1] == p:
            continue
        out.append(p)
    # remove revisits (keep the path simple)
    seen = {}
    simple = []
    for p in out:
        if p in seen:
            k = seen[p]
            for q in simple[k + 1:]:
                del seen[q]
            del simple[k + 1:]
        else:
            seen[p] = len(simple)
            simple.append(p)
    return np.array(simple, dtype=np.int64)


def lattice2d(nx, ny, Lx=1.0, Ly=1.0, jitter=0.2, seed=1, curves=(), hilbert=True, perturb=0.0,
              rotate_tds=True, flat_patch=0, flat_patch_perturb=0.0):
    """(nx x ny) quads, 2 triangles each.  `curves` = list of (points[M,2], closed) interface polylines.
    flat_patch = P > 0: the P x P lattice vertices in the lower-left corner keep their exact lattice positions (all quads
    there are cocircular, all rows collinear: the exact-predicate fallback has work to do, like config 4); with
    flat_patch_perturb = e > 0 every coordinate of the patch is moved by U(-e, e), so that signs are decided at the last bit."""
    m = Mesh2D()
    hx, hy = Lx / nx, Ly / ny
    m.h = min(hx, hy)
    I, J = np.meshgrid(np.arange(nx + 1, dtype=np.int64), np.arange(ny + 1, dtype=np.int64), indexing="xy")
    I, J = I.ravel(), J.ravel()
    lid = J * (nx + 1) + I                                   # persistent id = lattice row-major index
    nv = lid.shape[0]
    x = I.astype(np.float64) * hx
    y = J.astype(np.float64) * hy
    boundary = (I == 0) | (I == nx) | (J == 0) | (J == ny)

    is_if = np.zeros(nv, bool)
    forced = {}                                              # quad id -> 0 ('/') or 1 ('\')
    seg_l = []
    for pts, closed in curves:
        path = trace_curve(np.asarray(pts, float), hx, hy, nx, ny, closed)
        pid = path[:, 1] * (nx + 1) + path[:, 0]
        is_if[pid] = True
        a = path
        b = np.roll(path, -1, axis=0) if closed else path[1:]
        if not closed:
            a = path[:-1]
        di, dj = b[:, 0] - a[:, 0], b[:, 1] - a[:, 1]
        for k in np.nonzero((di != 0) & (dj != 0))[0]:
            qi, qj = min(a[k, 0], b[k, 0]), min(a[k, 1], b[k, 1])
            forced[qj * nx + qi] = 0 if di[k] * dj[k] > 0 else 1
        seg_l.append(np.stack([a[:, 1] * (nx + 1) + a[:, 0], b[:, 1] * (nx + 1) + b[:, 0]], axis=1))

    if jitter > 0:
        free = ~(boundary | is_if)
        if flat_patch > 0:
            free &= ~((I < flat_patch) & (J < flat_patch))
        x = x + np.where(free, (2.0 * uniform01(seed, lid, 0) - 1.0) * (jitter * hx), 0.0)
        y = y + np.where(free, (2.0 * uniform01(seed, lid, 1) - 1.0) * (jitter * hy), 0.0)
    if flat_patch > 0 and flat_patch_perturb > 0:
        inp = (I < flat_patch) & (J < flat_patch) & ~boundary
        x = x + np.where(inp, (2.0 * uniform01(seed, lid, 4) - 1.0) * flat_patch_perturb, 0.0)
        y = y + np.where(inp, (2.0 * uniform01(seed, lid, 5) - 1.0) * flat_patch_perturb, 0.0)
    if perturb > 0:                                          # config 4: every coordinate +- perturb
        x = x + (2.0 * uniform01(seed, lid, 2) - 1.0) * perturb
        y = y + (2.0 * uniform01(seed, lid, 3) - 1.0) * perturb

    QI, QJ = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(ny, dtype=np.int64), indexing="xy")
    QI, QJ = QI.ravel(), QJ.ravel()
    v00 = QJ * (nx + 1) + QI
    v10, v01, v11 = v00 + 1, v00 + nx + 1, v00 + nx + 2
    d_slash = (x[v11] - x[v00]) ** 2 + (y[v11] - y[v00]) ** 2
    d_back = (x[v01] - x[v10]) ** 2 + (y[v01] - y[v10]) ** 2
    diag = (d_back < d_slash).astype(np.int8)                # ties -> '/'
    if forced:
        fk = np.fromiter(forced.keys(), dtype=np.int64)
        diag[fk] = np.fromiter(forced.values(), dtype=np.int8)
    s = diag == 0
    t0 = np.where(s[:, None], np.stack([v00, v10, v11], 1), np.stack([v00, v10, v01], 1))
    t1 = np.where(s[:, None], np.stack([v00, v11, v01], 1), np.stack([v10, v11, v01], 1))
    cells = np.empty((2 * nx * ny, 3), np.int64)
    cells[0::2], cells[1::2] = t0, t1
    if rotate_tds:                                           # TDS start vertex is arbitrary: rotate pseudo-randomly
        rot = (splitmix64(np.arange(cells.shape[0], dtype=np.uint64) + _U64(seed) * _U64(7919)) % _U64(3)).astype(np.int64)
        idx = (np.arange(3)[None, :] + rot[:, None]) % 3
        cells = np.take_along_axis(cells, idx, axis=1)

    if hilbert:
        order = int(np.ceil(np.log2(max(nx, ny) + 1)))
        vrank = np.empty(nv, np.int64)
        vorder = np.argsort(hilbert_key(order, I, J), kind="stable")
        vrank[vorder] = np.arange(nv)
        ckey = hilbert_key(order, np.repeat(QI, 2), np.repeat(QJ, 2)) * 2 + np.tile(np.array([0, 1]), nx * ny)
        corder = np.argsort(ckey, kind="stable")
        cells = cells[corder]
    else:
        vorder = np.arange(nv)
        vrank = np.arange(nv)
    m.vid = lid[vorder]
    m.xy = np.ascontiguousarray(np.stack([x[vorder], y[vorder]], axis=1))
    m.lat = np.stack([I[vorder], J[vorder]], axis=1)
    m.vflags = (is_if[vorder].astype(np.uint8) | (boundary[vorder].astype(np.uint8) << 1)).astype(np.uint8)
    m.perm = pack_perm(cells)                                # cells still hold persistent ids here
    m.cells = np.ascontiguousarray(vrank[cells].astype(np.int32))
    m.edges, m.nb, m.edge_cell = build_edges(m.cells, nv)
    if seg_l:
        sg = np.concatenate(seg_l, axis=0)
        sg = np.where((sg[:, 0] < sg[:, 1])[:, None], sg, sg[:, ::-1])   # lower persistent id first
        m.segs = np.ascontiguousarray(vrank[sg].astype(np.int32))
    else:
        m.segs = np.zeros((0, 2), np.int32)
    return m


# ---------------------------------------------------------------------------------------------------------
# The named configurations (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------------------
def _line(y0, Lx, n):
    xs = np.linspace(0.0, Lx, n)
    return np.stack([xs, np.full_like(xs, y0)], 1)


def config1(n=100, seed=1):
    """~20k triangles on the unit square, straight interface y = 0.5."""
    m = lattice2d(n, n, 1.0, 1.0, jitter=0.2, seed=seed, curves=[(_line(0.5, 1.0, 16 * n + 1), False)])
    return m


def config2(n=708, seed=2):
    """~1M triangles, closed circular interface c=(.5,.5), R=.25."""
    t = np.linspace(0.0, 2 * np.pi, 64 * n, endpoint=False)
    circ = np.stack([0.5 + 0.25 * np.cos(t), 0.5 + 0.25 * np.sin(t)], 1)
    return lattice2d(n, n, 1.0, 1.0, jitter=0.2, seed=seed, curves=[(circ, True)])


def config3(nx=4096, ny=2048, seed=3, flat_patch=0, flat_patch_perturb=0.0):
    """16.8M triangles on [0,2]x[0,1], Hilbert-sorted, 4 sinusoidal interface polylines."""
    xs = np.linspace(0.0, 2.0, 16 * nx + 1)
    curves = [(np.stack([xs, 0.2 * k + 0.05 * np.sin(2 * np.pi * xs)], 1), False) for k in (1, 2, 3, 4)]
    return lattice2d(nx, ny, 2.0, 1.0, jitter=0.2, seed=seed, curves=curves, flat_patch=flat_patch, flat_patch_perturb=flat_patch_perturb)


def config4(n=1024, perturb=1e-15, seed=4):
    """Near-degenerate cocircular/collinear lattice (exact lattice when perturb=0)."""
    return lattice2d(n, n, 1.0, 1.0, jitter=0.0, seed=seed, perturb=perturb, hilbert=True)


def shifts_uniform(m, amp, seed):
    """U(-amp, amp)^dim per vertex, zero on the domain boundary."""
    dim = m.xy.shape[1] if m.dim == 2 else 3
    x = m.xy if m.dim == 2 else m.xyz
    s = np.empty_like(x)
    for k in range(dim):
        s[:, k] = (2.0 * uniform01(seed, m.vid, k) - 1.0) * amp
    s[(m.vflags & 2) != 0] = 0.0
    return s


def shifts_rotation(m, dtheta, center=(0.5, 0.5), r_in=0.35, r_out=0.5):
    """Rigid rotation by dtheta about `center` inside r_in, smoothly cut off to zero at r_out
    (the rotating field of dune/mmesh/test/examples/movement.cc:49-59 made boundary-safe)."""
    dx = m.xy[:, 0] - center[0]
    dy = m.xy[:, 1] - center[1]
    r = np.hypot(dx, dy)
    t = np.clip((r_out - r) / (r_out - r_in), 0.0, 1.0)
    w = t * t * (3.0 - 2.0 * t)
    c, s = np.cos(dtheta * w), np.sin(dtheta * w)
    out = np.stack([c * dx - s * dy - dx, s * dx + c * dy - dy], 1)
    return np.ascontiguousarray(out)


def pairs_self_and_neighbours(m, xy_old):
    """Projection stress pairs: every cell is 'new'; its old children are its own
    pre-move copy and its face neighbours' copies.  Old corners are cached in id order through
    AffineGeometry::corner (cachingentity.hh:83-89)."""
    nc = m.nc
    nbv = m.nb >= 0
    cnt = 1 + nbv.sum(axis=1)
    new_off = np.zeros(nc + 1, np.int64)
    np.cumsum(cnt, out=new_off[1:])
    P = int(new_off[-1])
    old_idx = np.empty(P, np.int32)
    pos = new_off[:-1].copy()
    old_idx[pos] = np.arange(nc, dtype=np.int32)
    pos += 1
    for k in range(3):
        sel = nbv[:, k]
        old_idx[pos[sel]] = m.nb[sel, k]
        pos[sel] += 1
    old_xy = cached_corners(m, xy_old, old_idx)
    return new_off, np.arange(nc, dtype=np.int32), old_xy, old_idx


def cached_corners(m, xy_old, cell_idx):
    """CachingEntity::vertex_[i] = geo.corner(i) of the id-ordered AffineGeometry (origin + J^T e_i)."""
    c = m.cells[cell_idx]
    p = m.perm[cell_idx]
    k = [np.take_along_axis(c, ((p >> (2 * i)) & 3)[:, None].astype(np.int64), axis=1)[:, 0] for i in range(3)]
    c0, c1, c2 = xy_old[k[0]], xy_old[k[1]], xy_old[k[2]]
    j0, j1 = c1 - c0, c2 - c0
    out = np.empty((c.shape[0], 6))
    out[:, 0:2] = c0
    out[:, 2:4] = (c0 + 1.0 * j0) + 0.0 * j1
    out[:, 4:6] = (c0 + 0.0 * j0) + 1.0 * j1
    return out


def flip_case(n, jitter=0.2, seed=7):
    """Exactly conservative projection set-up at any size (vectorised): old mesh = jittered n x n lattice, new mesh = the
    same vertices with the OTHER diagonal in every quad; the connected component of a new triangle is its quad
    (two old children, connectedcomponent.hh:104-127).  Returns (new-mesh snapshot arrays, pair CSR, old areas/centroids)."""
    old = lattice2d(n, n, 1.0, 1.0, jitter=jitter, seed=seed, hilbert=False, rotate_tds=False)
    xy = old.xy
    q = np.arange(n * n, dtype=np.int64)
    qi, qj = q % n, q // n
    v00 = qj * (n + 1) + qi
    v10, v01, v11 = v00 + 1, v00 + n + 1, v00 + n + 2
    slash = old.cells[0::2, 2] == v11                       # old diagonal v00-v11 ('/'): T0 = (v00, v10, v11)
    # new triangles use the other diagonal, ccw
    n0 = np.where(slash[:, None], np.stack([v00, v10, v01], 1), np.stack([v00, v10, v11], 1))
    n1 = np.where(slash[:, None], np.stack([v10, v11, v01], 1), np.stack([v00, v11, v01], 1))
    cells = np.empty((2 * n * n, 3), np.int64)
    cells[0::2], cells[1::2] = n0, n1
    perm = pack_perm(old.vid[cells])
    nnew = cells.shape[0]
    new_off = np.arange(0, 2 * nnew + 1, 2, dtype=np.int64)
    old_idx = np.repeat(np.stack([2 * q, 2 * q + 1], 1), 2, axis=0).reshape(-1).astype(np.int32)   # both old children per new cell
    old_xy = cached_corners(old, xy, old_idx)
    po = xy[old.cells]
    old_vol = np.abs(cross2(po[:, 1] - po[:, 0], po[:, 2] - po[:, 0])) * 0.5
    return dict(xy=xy, cells=cells.astype(np.int32), perm=perm, new_off=new_off, new_cell=np.arange(nnew, dtype=np.int32),
                old_xy=old_xy, old_idx=old_idx, n_old=old.nc, old_vol=old_vol, old_centroid=po.mean(axis=1))


# ---------------------------------------------------------------------------------------------------------
# 3-D (config 5): Kuhn tetrahedra on an n^3 lattice, planar interface z = 0.5
# ---------------------------------------------------------------------------------------------------------
class Mesh3D:
    dim = 3

    def __init__(self):
        self.xyz = self.vid = self.vflags = self.cells = self.perm = self.tris = None
        self.h = None

    @property
    def nv(self):
        return self.xyz.shape[0]

    @property
    def nc(self):
        return self.cells.shape[0]


_KUHN = [(0, 1, 3, 7), (0, 1, 5, 7), (0, 2, 3, 7), (0, 2, 6, 7), (0, 4, 5, 7), (0, 4, 6, 7)]


def lattice3d(n, jitter=0.15, seed=5, interface_plane=True):
    m = Mesh3D()
    h = 1.0 / n
    m.h = h
    g = np.arange(n + 1, dtype=np.int64)
    K, J, I = np.meshgrid(g, g, g, indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    lid = (K * (n + 1) + J) * (n + 1) + I
    nv = lid.shape[0]
    kz = n // 2
    boundary = (I == 0) | (I == n) | (J == 0) | (J == n) | (K == 0) | (K == n)
    is_if = (K == kz) if interface_plane else np.zeros(nv, bool)
    free = ~(boundary | is_if)
    xyz = np.stack([I * h, J * h, K * h], 1).astype(np.float64)
    for c in range(3):
        xyz[:, c] += np.where(free, (2.0 * uniform01(seed, lid, c) - 1.0) * (jitter * h), 0.0)
    q = np.arange(n, dtype=np.int64)
    QK, QJ, QI = np.meshgrid(q, q, q, indexing="ij")
    QI, QJ, QK = QI.ravel(), QJ.ravel(), QK.ravel()
    base = (QK * (n + 1) + QJ) * (n + 1) + QI
    corner = [base + ((b >> 2) & 1) * (n + 1) * (n + 1) + ((b >> 1) & 1) * (n + 1) + (b & 1) for b in range(8)]
    tets = []
    for t in _KUHN:
        tets.append(np.stack([corner[t[0]], corner[t[1]], corner[t[2]], corner[t[3]]], 1))
    cells = np.stack(tets, 1).reshape(-1, 4)
    # make every tet positively oriented (CGAL TDS convention): swap last two vertices when negative
    p = xyz[cells]
    a, b, c = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]
    det = np.einsum("ij,ij->i", a, np.cross(b, c))
    neg = det < 0
    cells[neg] = cells[neg][:, [0, 1, 3, 2]]
    # Morton-ish locality: cells already cube-major; vertices stay lattice-ordered
    m.vid = lid
    m.xyz = np.ascontiguousarray(xyz)
    m.vflags = (is_if.astype(np.uint8) | (boundary.astype(np.uint8) << 1)).astype(np.uint8)
    m.perm = pack_perm(cells)
    m.cells = np.ascontiguousarray(cells.astype(np.int32))
    if interface_plane:
        s = np.arange(n, dtype=np.int64)
        SJ, SI = np.meshgrid(s, s, indexing="ij")
        SI, SJ = SI.ravel(), SJ.ravel()
        b0 = (kz * (n + 1) + SJ) * (n + 1) + SI
        # Kuhn faces on a z-plane: diagonal (0,0)-(1,1)
        ta = np.stack([b0, b0 + 1, b0 + n + 2], 1)
        tb = np.stack([b0, b0 + n + 1, b0 + n + 2], 1)
        m.tris = np.ascontiguousarray(np.concatenate([ta, tb]).astype(np.int32))
    else:
        m.tris = np.zeros((0, 3), np.int32)
    return m
