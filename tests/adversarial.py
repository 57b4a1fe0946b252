"""
Adversarial geometry for the restated Shapely predicates, decided by RATIONAL arithmetic.

The collision path (`Polygon.intersects` behind state/utils.py:10-49 / utils.py:28-62) and the sensor's
radius query (`Point.buffer(r)` + `vectorized.contains`, state/state.py:352-372) are restated, not
linked: GEOS is absent.  What pins the restatements beyond the reference's few known-answer tests is
this set: configurations on the knife edge -- boxes touching at corners, along collinear edges, corner
on edge, each also moved by one ulp either way; zero-area and zero-length boxes; identical boxes; query
points on the vertices / edges of the 64-gon -- whose answers are computed here exactly, with
`fractions.Fraction`, from the same fp64 corner / vertex coordinates the engines compute.

Boxes are axis-aligned (heading 0) so that every implementation derives bit-identical corners (no
transcendental rounding in between); rotated knife edges cannot be pinned this way and are covered
only statistically by the random pairs of tests/golden/unit.npz.
"""
from __future__ import annotations

from fractions import Fraction as F
from typing import List, Tuple

import numpy as np


# ------------------------------------------------------------------------------ exact predicates
def _orient(a, b, c) -> int:
    d = (F(b[0]) - F(a[0])) * (F(c[1]) - F(a[1])) - (F(b[1]) - F(a[1])) * (F(c[0]) - F(a[0]))
    return (d > 0) - (d < 0)


def _on_segment(a, b, p) -> bool:
    return (_orient(a, b, p) == 0 and min(a[0], b[0]) <= p[0] <= max(a[0], b[0])
            and min(a[1], b[1]) <= p[1] <= max(a[1], b[1]))


def _segments_meet(a, b, c, d) -> bool:
    """Closed segments ab and cd share a point (exact)."""
    o1, o2, o3, o4 = _orient(a, b, c), _orient(a, b, d), _orient(c, d, a), _orient(c, d, b)
    if o1 * o2 < 0 and o3 * o4 < 0:
        return True
    return _on_segment(a, b, c) or _on_segment(a, b, d) or _on_segment(c, d, a) or _on_segment(c, d, b)


def _inside_closed(poly, p) -> bool:
    """p in the closed convex polygon `poly` of non-zero area (exact)."""
    signs = [_orient(poly[k], poly[(k + 1) % len(poly)], p) for k in range(len(poly))]
    return not (any(s > 0 for s in signs) and any(s < 0 for s in signs))


def _area2(poly) -> F:
    return sum(F(poly[k][0]) * F(poly[(k + 1) % 4][1]) - F(poly[(k + 1) % 4][0]) * F(poly[k][1]) for k in range(4))


def quads_intersect_exact(qa: np.ndarray, qb: np.ndarray) -> bool:
    """
    Closed-set intersection of the quads with corner rows qa, qb (4 x 2, fp64), exactly: their rings
    meet, or one lies inside the other (only a quad with area can contain something).  Quads with
    bit-identical corner arrays are excluded -- the reference drops them (`g != g_prime`, utils.py:58).
    """
    if np.array_equal(qa, qb):
        return False
    A = [tuple(map(float, r)) for r in qa]
    B = [tuple(map(float, r)) for r in qb]
    for i in range(4):
        for j in range(4):
            if _segments_meet(A[i], A[(i + 1) % 4], B[j], B[(j + 1) % 4]):
                return True
    if _area2(A) != 0 and _inside_closed(A, B[0]):
        return True
    if _area2(B) != 0 and _inside_closed(B, A[0]):
        return True
    return False


def corners(pose3, box) -> np.ndarray:
    """Entity.get_bounding_box_points for heading 0 (entity/base.py:100-138: R = identity, exact)."""
    x, y, h = pose3
    assert h == 0.0
    w, l, cx, cy = box
    pts = np.array([[cx - 0.5 * l, cy + 0.5 * w], [cx + 0.5 * l, cy + 0.5 * w],
                    [cx + 0.5 * l, cy - 0.5 * w], [cx - 0.5 * l, cy - 0.5 * w]])
    # points @ R with c = 1, s = 0: p0 * 1 + p1 * -0.0 and p0 * 0 + p1 * 1 are exact
    return np.array([x, y]) + pts


# ------------------------------------------------------------------------------ box pairs
def box_pair_cases() -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray, List[str]]:
    """(pose_a [n,3], box_a [n,4], pose_b, box_b, expected [n] bool, labels)."""
    up = lambda v: float(np.nextafter(v, np.inf))  # noqa: E731
    dn = lambda v: float(np.nextafter(v, -np.inf))  # noqa: E731
    cases = []

    def add(label, pa, ba, pb, bb):
        cases.append((label, (pa[0], pa[1], 0.0), tuple(map(float, ba)), (pb[0], pb[1], 0.0), tuple(map(float, bb))))

    A = (2.0, 4.0, 0.0, 0.0)       # width 2, length 4, centred
    off = (2.0, 4.2, 1.37, 0.0)    # the reference's car1: box not centred on the pose
    for name, box in (("centred", A), ("car1", off)):
        w, l = box[0], box[1]
        for tag, f in (("", lambda v: v), ("+ulp", up), ("-ulp", dn)):
            add(f"{name}: corner to corner{tag}", (0.0, 0.0), box, (f(l), f(w)), box)
            add(f"{name}: corner to corner, x only{tag}", (0.0, 0.0), box, (f(l), w), box)
            add(f"{name}: edge to edge (end faces){tag}", (0.0, 0.0), box, (f(l), 0.25), box)
            add(f"{name}: edge to edge (sides), collinear part{tag}", (0.0, 0.0), box, (1.5, f(w)), box)
            add(f"{name}: corner on edge interior{tag}", (0.0, 0.0), box, (f(l), 0.5 * w), (1.0, 1.0, 0.0, 0.0))
            add(f"{name}: small box inside, touching the wall{tag}", (0.0, 0.0), box,
                (f(0.5 * l - 0.25 + box[2]), 0.0), (0.5, 0.5, 0.0, 0.0))
        add(f"{name}: strictly inside", (0.0, 0.0), box, (box[2], 0.0), (0.5, 0.5, 0.0, 0.0))
        add(f"{name}: identical", (3.0, -1.0), box, (3.0, -1.0), box)
        add(f"{name}: identical position, other width", (3.0, -1.0), box, (3.0, -1.0), (w + 0.5, l, box[2], box[3]))
        add(f"{name}: identical + ulp", (3.0, -1.0), box, (up(3.0), -1.0), box)
        add(f"{name}: far apart", (0.0, 0.0), box, (50.0, 50.0), box)
    # large coordinates: one ulp is 2^-42 of the magnitude
    big = 4096.0
    for tag, f in (("", lambda v: v), ("+ulp", up), ("-ulp", dn)):
        add(f"far from the origin: end faces touching{tag}", (big, big), A, (f(big + 4.0), big), A)
    # zero-area and zero-length boxes
    seg_h = (0.0, 4.0, 0.0, 0.0)   # width 0: a segment along x
    seg_v = (2.0, 0.0, 0.0, 0.0)   # length 0: a segment along y
    pt = (0.0, 0.0, 0.0, 0.0)      # a point
    add("segment through a box", (0.0, 0.0), A, (1.0, 0.0), seg_h)
    add("segment inside a box", (0.0, 0.0), A, (0.0, 0.0), (0.0, 1.0, 0.0, 0.0))
    add("segment on a box's side, collinear", (0.0, 0.0), A, (1.0, 1.0), seg_h)
    add("segment on a box's side, collinear +ulp", (0.0, 0.0), A, (1.0, up(1.0)), seg_h)
    add("segment touching a box's end face with its end point", (0.0, 0.0), A, (4.0, 0.0), seg_h)
    add("segment short of a box by an ulp", (0.0, 0.0), A, (up(4.0), 0.0), seg_h)
    add("segment beside a box", (0.0, 0.0), A, (0.0, 3.0), seg_h)
    add("segments crossing", (0.0, 0.0), seg_h, (0.5, 0.25), seg_v)
    add("segments touching at an end point (T)", (0.0, 0.0), seg_h, (1.0, 1.0), seg_v)
    add("segments touching at an end point (T) +ulp", (0.0, 0.0), seg_h, (1.0, up(1.0)), seg_v)
    add("segments collinear, overlapping", (0.0, 0.0), seg_h, (3.0, 0.0), seg_h)
    add("segments collinear, end to end", (0.0, 0.0), seg_h, (4.0, 0.0), seg_h)
    add("segments collinear, apart", (0.0, 0.0), seg_h, (up(4.0), 0.0), seg_h)
    add("segments parallel", (0.0, 0.0), seg_h, (0.0, 0.5), seg_h)
    add("segments in overlapping AABBs that miss each other", (0.0, 0.0), seg_v, (1.0, 0.0), (0.0, 1.0, 0.0, 0.0))
    add("point on a box corner", (0.0, 0.0), A, (2.0, 1.0), pt)
    add("point on a box side", (0.0, 0.0), A, (0.5, 1.0), pt)
    add("point inside a box", (0.0, 0.0), A, (0.5, 0.5), pt)
    add("point an ulp outside a box", (0.0, 0.0), A, (0.5, up(1.0)), pt)
    add("point on a segment", (0.0, 0.0), seg_h, (1.0, 0.0), pt)
    add("point beside a segment", (0.0, 0.0), seg_h, (1.0, up(0.0)), pt)
    add("two identical points", (1.0, 1.0), pt, (1.0, 1.0), pt)
    add("two distinct points", (1.0, 1.0), pt, (up(1.0), 1.0), pt)
    labels = [c[0] for c in cases]
    pa = np.array([c[1] for c in cases])
    ba = np.array([c[2] for c in cases])
    pb = np.array([c[3] for c in cases])
    bb = np.array([c[4] for c in cases])
    want = np.array([quads_intersect_exact(corners(pa[k], ba[k]), corners(pb[k], bb[k])) for k in range(len(cases))])
    return pa, ba, pb, bb, want, labels


# ------------------------------------------------------------------------------ the 64-gon
def ngon_vertices(x: float, y: float, r: float) -> np.ndarray:
    """Vertices of GEOS' Point(x, y).buffer(r) as the engines compute them: x + r * cos(-k * 2 pi / 64), ..."""
    import math

    inc = (2.0 * math.pi) / 64
    return np.array([[x + r * math.cos(0.0 + -1.0 * k * inc), y + r * math.sin(0.0 + -1.0 * k * inc)] for k in range(64)])


def in_ngon_exact(v: np.ndarray, qx: float, qy: float) -> bool:
    """(qx, qy) strictly inside the clockwise polygon v (exact)."""
    return all(_orient(tuple(v[k]), tuple(v[(k + 1) % 64]), (qx, qy)) < 0 for k in range(64))


def ngon_cases():
    """(x, y, r, qx, qy, expected, label) rows: query points on / next to vertices and edges of the 64-gon."""
    up = lambda v: float(np.nextafter(v, np.inf))  # noqa: E731
    dn = lambda v: float(np.nextafter(v, -np.inf))  # noqa: E731
    rows = []
    for (x, y, r) in ((0.0, 0.0, 1.0), (12.5, -3.25, 1.0), (100.0, 200.0, 3.0), (-7.0, 0.5, 0.125)):
        v = ngon_vertices(x, y, r)
        for k in (0, 1, 7, 16, 31, 32, 48, 63):
            vx, vy = map(float, v[k])
            mx, my = 0.5 * (vx + float(v[(k + 1) % 64, 0])), 0.5 * (vy + float(v[(k + 1) % 64, 1]))
            for label, (qx, qy) in (("vertex", (vx, vy)), ("vertex, x-ulp", (dn(vx) if vx > x else up(vx), vy)),
                                    ("vertex, x+ulp", (up(vx) if vx > x else dn(vx), vy)),
                                    ("edge midpoint", (mx, my)),
                                    ("towards the centre of the edge midpoint", (x + (mx - x) * (1 - 2 ** -40), y + (my - y) * (1 - 2 ** -40))),
                                    ("beyond the edge midpoint", (x + (mx - x) * (1 + 2 ** -40), y + (my - y) * (1 + 2 ** -40)))):
                rows.append((x, y, r, qx, qy, in_ngon_exact(v, qx, qy), f"r={r} k={k} {label}"))
        rows.append((x, y, r, x, y, True, "centre"))
        rows.append((x, y, r, x + 0.999 * r, y, True, "inside, near vertex 0"))
        rows.append((x, y, r, x + r * 0.9988 * np.cos(np.pi / 64), y - r * 0.9988 * np.sin(np.pi / 64),
                     in_ngon_exact(v, x + r * 0.9988 * np.cos(np.pi / 64), y - r * 0.9988 * np.sin(np.pi / 64)),
                     "on the apothem band"))
    return rows
