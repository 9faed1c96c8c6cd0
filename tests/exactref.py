"""Exact (python fractions.Fraction) references used to pin the oracle: doubles are exact rationals."""
from fractions import Fraction as F


def sgn(x):
    return (x > 0) - (x < 0)


def fr(a):
    return [F(float(v)) for v in a]


def orient2d(p, q, r):
    p, q, r = fr(p), fr(q), fr(r)
    return sgn((q[0] - p[0]) * (r[1] - p[1]) - (r[0] - p[0]) * (q[1] - p[1]))


def incircle(p, q, r, t):
    """CGAL/predicates/kernel_ftC2.h:477-499 evaluated over the rationals."""
    p, q, r, t = fr(p), fr(q), fr(r), fr(t)
    qpx, qpy = q[0] - p[0], q[1] - p[1]
    rpx, rpy = r[0] - p[0], r[1] - p[1]
    tpx, tpy = t[0] - p[0], t[1] - p[1]
    a00 = qpx * tpy - qpy * tpx
    a01 = tpx * (t[0] - q[0]) + tpy * (t[1] - q[1])
    a10 = qpx * rpy - qpy * rpx
    a11 = rpx * (r[0] - q[0]) + rpy * (r[1] - q[1])
    return sgn(a00 * a11 - a10 * a01)


def incircle_lifted(p, q, r, t):
    """Textbook 3x3 lifted determinant (independent formula): >0 iff t inside the ccw circle pqr."""
    p, q, r, t = fr(p), fr(q), fr(r), fr(t)
    rows = []
    for a in (p, q, r):
        dx, dy = a[0] - t[0], a[1] - t[1]
        rows.append((dx, dy, dx * dx + dy * dy))
    (a, b, c), (d, e, f), (g, h, i) = rows
    return sgn(a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g))


def orient3d(p, q, r, s):
    p, q, r, s = fr(p), fr(q), fr(r), fr(s)
    a = [[x[j] - p[j] for j in range(3)] for x in (q, r, s)]
    m01 = a[0][0] * a[1][1] - a[1][0] * a[0][1]
    m02 = a[0][0] * a[2][1] - a[2][0] * a[0][1]
    m12 = a[1][0] * a[2][1] - a[2][0] * a[1][1]
    return sgn(m01 * a[2][2] - m02 * a[1][2] + m12 * a[0][2])


def clip_area_exact(subj, clip):
    """Exact area of (convex ccw polygon subj) ∩ (convex ccw polygon clip) with rational Sutherland-Hodgman."""
    poly = [tuple(fr(p)) for p in subj]
    cl = [tuple(fr(p)) for p in clip]
    for i in range(len(cl)):
        a, b = cl[i], cl[(i + 1) % len(cl)]
        if not poly:
            break

        def side(p):
            return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])

        out = []
        for k in range(len(poly)):
            p, q = poly[k], poly[(k + 1) % len(poly)]
            sp, sq = side(p), side(q)
            if sp >= 0:
                out.append(p)
            if (sp > 0 and sq < 0) or (sp < 0 and sq > 0):
                t = sp / (sp - sq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        poly = out
    if len(poly) < 3:
        return F(0)
    s = F(0)
    for k in range(len(poly)):
        p, q = poly[k], poly[(k + 1) % len(poly)]
        s += p[0] * q[1] - q[0] * p[1]
    return s / 2
