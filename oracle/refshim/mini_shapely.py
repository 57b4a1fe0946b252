"""
TEST INFRASTRUCTURE ONLY -- restated subset of Shapely >= 2.0 / GEOS.

The reference (driskai/scenario_gym v0.3.1) calls Shapely at a handful of sites on
the rollout hot path.  Shapely/GEOS are not installable in this image, so this
module restates the *published semantics* of exactly the calls the reference
makes, so that the reference's own Python code can run unmodified in this
container and emit golden vectors (see ``oracle/gen_golden.py``).

Call sites restated (reference file:line):
  * ``Polygon(pts)``, ``STRtree(geoms).query(g, predicate="intersects")``,
    ``tree.geometries.take``, ``g != g_prime``     scenario_gym/utils.py:51-62,
                                                   scenario_gym/state/utils.py:38-48
  * ``Polygon.intersects(Polygon|LineString)``, ``Polygon.area``
                                                   metrics/rss/callback.py:186-196,317-328
  * ``Point(x, y).buffer(r)`` + ``shapely.vectorized.contains``
                                                   state/state.py:352-372
  * ``LineString(route).project(Point)``           pedestrian/agent.py:45,61

Semantics followed: GEOS ``intersects`` is closed-set (touching counts) and is
decided with robust orientation predicates on the fp64 coordinates; here every
orientation sign is decided exactly (fp64 filter, then rational arithmetic), so
the result is the exact closed-set answer for the given fp64 corners.
``Point.buffer(r)`` is GEOS' 64-gon (quad_segs=16, vertices generated clockwise
from angle 0).  ``contains`` is strict interior membership.  Geometry equality is
Shapely 2's structural equality (same type, identical coordinate sequence).

Parity note: GEOS itself is absent, so this restatement is pinned only by the
reference tests that exercise it (tests/test_utils.py:43-61 head-on boxes,
tests/test_state.py:74-100 radius query counts, tests/pedestrian/test_ped_sensor.py).
"""
from __future__ import annotations

import math
from fractions import Fraction
from typing import Iterable, List, Sequence

import numpy as np

_EPS = 2.0 ** -53
_CCW_ERRBOUND = (3.0 + 16.0 * _EPS) * _EPS


def orient_sign(ax, ay, bx, by, cx, cy) -> int:
    """Exact sign of the 2x2 determinant |b-a, c-a| for fp64 inputs."""
    detleft = (ax - cx) * (by - cy)
    detright = (ay - cy) * (bx - cx)
    det = detleft - detright
    detsum = abs(detleft) + abs(detright)
    if abs(det) > _CCW_ERRBOUND * detsum:
        return 1 if det > 0 else -1
    F = Fraction
    d = (F(ax) - F(cx)) * (F(by) - F(cy)) - (F(ay) - F(cy)) * (F(bx) - F(cx))
    return (d > 0) - (d < 0)


def _ring_orientation(pts: np.ndarray) -> int:
    """+1 for counter-clockwise, -1 for clockwise, 0 for degenerate (exact)."""
    F = Fraction
    a = F(0)
    n = len(pts)
    for i in range(n):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % n]
        a += F(float(x0)) * F(float(y1)) - F(float(x1)) * F(float(y0))
    return (a > 0) - (a < 0)


class _Coords:
    def __init__(self, arr: np.ndarray):
        self._arr = arr

    def __getitem__(self, idx):
        out = self._arr[idx]
        if out.ndim == 1:
            return tuple(out.tolist())
        return [tuple(r) for r in out.tolist()]

    def __len__(self):
        return len(self._arr)

    def __iter__(self):
        return iter([tuple(r) for r in self._arr.tolist()])

    @property
    def xy(self):
        return self._arr[:, 0].copy(), self._arr[:, 1].copy()


class BaseGeometry:
    """Common structural equality / hashing, as in Shapely 2."""

    _pts: np.ndarray

    def _key(self):
        return (type(self).__name__, self._pts.tobytes())

    def __eq__(self, other):
        if not isinstance(other, BaseGeometry):
            return NotImplemented
        return self._key() == other._key()

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    def __hash__(self):
        return hash(self._key())


class Point(BaseGeometry):
    def __init__(self, *args):
        if len(args) == 1:
            args = tuple(args[0])
        self.x = float(args[0])
        self.y = float(args[1])
        self._pts = np.array([[self.x, self.y]], dtype=np.float64)

    @property
    def xy(self):
        return np.array([self.x]), np.array([self.y])

    @property
    def area(self):
        return 0.0

    def buffer(self, r: float, quad_segs: int = 16) -> "Polygon":
        """GEOS point buffer: 4*quad_segs-gon, clockwise from angle 0."""
        n = 4 * quad_segs
        inc = (2.0 * math.pi) / n
        pts = []
        for i in range(n):
            ang = 0.0 + -1.0 * i * inc
            pts.append((self.x + r * math.cos(ang), self.y + r * math.sin(ang)))
        return Polygon(pts)


class LineString(BaseGeometry):
    def __init__(self, coords: Iterable):
        self._pts = np.array([tuple(map(float, c))[:2] for c in coords], dtype=np.float64)

    @property
    def coords(self):
        return _Coords(self._pts)

    @property
    def area(self):
        return 0.0

    @property
    def length(self):
        return float(np.linalg.norm(np.diff(self._pts, axis=0), axis=1).sum())

    def project(self, other: Point) -> float:
        """Distance along the line to the point nearest to ``other``."""
        px, py = other.x, other.y
        best_d = math.inf
        best_s = 0.0
        s0 = 0.0
        for i in range(len(self._pts) - 1):
            ax, ay = self._pts[i]
            bx, by = self._pts[i + 1]
            dx, dy = bx - ax, by - ay
            seg2 = dx * dx + dy * dy
            seglen = math.sqrt(seg2)
            if seg2 == 0.0:
                r = 0.0
            else:
                r = ((px - ax) * dx + (py - ay) * dy) / seg2
            if r <= 0.0:
                qx, qy, sl = ax, ay, 0.0
            elif r >= 1.0:
                qx, qy, sl = bx, by, seglen
            else:
                qx, qy = ax + r * dx, ay + r * dy
                sl = r * seglen
            d = math.hypot(px - qx, py - qy)
            if d < best_d:
                best_d = d
                best_s = s0 + sl
            s0 += seglen
        return best_s


def _open_ring(coords) -> np.ndarray:
    if isinstance(coords, LinearRing):
        return coords._pts
    pts = np.array([tuple(map(float, c))[:2] for c in coords], dtype=np.float64).reshape(-1, 2)
    if len(pts) > 1 and (pts[0] == pts[-1]).all():
        pts = pts[:-1]
    return pts


class LinearRing(BaseGeometry):
    def __init__(self, coords: Iterable):
        self._pts = _open_ring(coords)

    @property
    def coords(self):
        return _Coords(np.concatenate([self._pts, self._pts[:1]], axis=0))


def _ring_side(pts: np.ndarray, px: float, py: float) -> int:
    """+1 strictly inside the ring, 0 on it, -1 outside: crossing parity with exact orientation signs."""
    inside = False
    n = len(pts)
    for k in range(n):
        ax, ay = pts[k]
        bx, by = pts[(k + 1) % n]
        straddles = (ay > py) != (by > py)
        in_box = min(ax, bx) <= px <= max(ax, bx) and min(ay, by) <= py <= max(ay, by)
        if not straddles and not in_box:
            continue
        o = orient_sign(ax, ay, bx, by, px, py)
        if o == 0 and in_box:
            return 0
        if straddles and ((o > 0) == (by > ay)):
            inside = not inside
    return 1 if inside else -1


class Polygon(BaseGeometry):
    """Polygon with optional holes.  Box-to-box predicates assume convex shells; point membership
    (`contains`, road-network surfaces) is general."""

    is_valid = True

    def __init__(self, shell: Iterable, holes: Iterable = None):
        self._pts = _open_ring(shell)
        self._holes = [_open_ring(h) for h in (holes or [])]
        self._orient = None

    @property
    def interiors(self):
        return [LinearRing(h) for h in self._holes]

    def point_side(self, px: float, py: float) -> int:
        """+1 interior, 0 boundary, -1 exterior."""
        s = _ring_side(self._pts, px, py)
        if s <= 0:
            return s
        for h in self._holes:
            hs = _ring_side(h, px, py)
            if hs == 0:
                return 0
            if hs > 0:
                return -1
        return 1

    def rings(self):
        return [self._pts] + self._holes

    @property
    def exterior(self):
        closed = np.concatenate([self._pts, self._pts[:1]], axis=0)

        class _Ring:
            coords = _Coords(closed)

        return _Ring()

    @property
    def area(self) -> float:
        def shoelace(p):
            x, y = p[:, 0], p[:, 1]
            return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(np.roll(x, -1), y)))

        return shoelace(self._pts) - sum(shoelace(h) for h in getattr(self, "_holes", []))

    @property
    def centroid(self) -> Point:
        x, y = self._pts[:, 0], self._pts[:, 1]
        xn, yn = np.roll(x, -1), np.roll(y, -1)
        cr = x * yn - xn * y
        a = cr.sum() / 2.0
        if a == 0.0:
            return Point(float(x.mean()), float(y.mean()))
        return Point(float(((x + xn) * cr).sum() / (6 * a)), float(((y + yn) * cr).sum() / (6 * a)))

    @property
    def bounds(self):
        return (
            float(self._pts[:, 0].min()),
            float(self._pts[:, 1].min()),
            float(self._pts[:, 0].max()),
            float(self._pts[:, 1].max()),
        )

    def orientation(self) -> int:
        if self._orient is None:
            self._orient = _ring_orientation(self._pts)
        return self._orient

    # -- predicates ---------------------------------------------------------
    def _edge_separates(self, k: int, pts: Sequence) -> bool:
        """True if every point of pts is strictly outside edge k of self."""
        n = len(self._pts)
        ax, ay = self._pts[k]
        bx, by = self._pts[(k + 1) % n]
        o = self.orientation()
        for (px, py) in pts:
            s = orient_sign(ax, ay, bx, by, px, py)
            # outside of a CCW ring is the right side (negative orientation)
            if s * o >= 0:
                return False
        return True

    def intersects(self, other) -> bool:
        """Closed-set intersection of convex shapes (exact)."""
        if isinstance(other, Polygon):
            if self.orientation() == 0 or other.orientation() == 0:
                raise NotImplementedError("degenerate polygon")
            for k in range(len(self._pts)):
                if self._edge_separates(k, other._pts):
                    return False
            for k in range(len(other._pts)):
                if other._edge_separates(k, self._pts):
                    return False
            return True
        if isinstance(other, LineString):
            if len(other._pts) != 2:
                raise NotImplementedError("only single segments")
            if self.orientation() == 0:
                raise NotImplementedError("degenerate polygon")
            for k in range(len(self._pts)):
                if self._edge_separates(k, other._pts):
                    return False
            (ax, ay), (bx, by) = other._pts
            if ax == bx and ay == by:
                return True  # point not outside any edge => inside/on
            signs = [orient_sign(ax, ay, bx, by, px, py) for (px, py) in self._pts]
            if all(s > 0 for s in signs) or all(s < 0 for s in signs):
                return False
            return True
        if isinstance(other, Point):
            return not any(
                self._edge_separates(k, other._pts) for k in range(len(self._pts))
            )
        raise NotImplementedError(type(other))

    def contains(self, other) -> bool:
        """Strict interior membership for points."""
        if isinstance(other, Point):
            return self.point_side(float(other.x), float(other.y)) > 0
        raise NotImplementedError(type(other))

    def intersection(self, other):
        raise NotImplementedError(
            "Polygon.intersection is not restated (reference path "
            "metrics/collision.py:91 is unreachable without AttributeError)"
        )


class MultiPolygon(BaseGeometry):
    def __init__(self, polygons: Iterable = ()):
        self.geoms = list(polygons)
        self._pts = (
            np.concatenate([g._pts for g in self.geoms], axis=0)
            if self.geoms
            else np.zeros((0, 2))
        )

    @property
    def area(self) -> float:
        return float(sum(g.area for g in self.geoms))

    def contains(self, other) -> bool:
        return any(g.contains(other) for g in self.geoms)


def contains(area, xs, ys) -> np.ndarray:
    """shapely.vectorized.contains for a convex polygon: strict interior."""
    xs = np.atleast_1d(np.asarray(xs, dtype=np.float64))
    ys = np.atleast_1d(np.asarray(ys, dtype=np.float64))
    if isinstance(area, MultiPolygon):
        out = np.zeros(xs.shape, dtype=bool)
        for g in area.geoms:
            out |= contains(g, xs, ys)
        return out
    o = area.orientation()
    n = len(area._pts)
    out = np.ones(xs.shape, dtype=bool)
    for idx in range(len(xs)):
        px, py = float(xs[idx]), float(ys[idx])
        for k in range(n):
            ax, ay = area._pts[k]
            bx, by = area._pts[(k + 1) % n]
            if orient_sign(ax, ay, bx, by, px, py) * o <= 0:
                out[idx] = False
                break
    return out


class _GeomArray:
    def __init__(self, geoms: List):
        self._g = list(geoms)

    def take(self, idx):
        return [self._g[i] for i in idx]

    def __getitem__(self, i):
        return self._g[i]

    def __len__(self):
        return len(self._g)


class STRtree:
    """Brute-force stand-in: envelope filter then exact predicate."""

    def __init__(self, geoms: Iterable):
        self.geometries = _GeomArray(list(geoms))

    def query(self, geometry, predicate=None) -> np.ndarray:
        hits = []
        b = geometry.bounds
        for i, g in enumerate(self.geometries._g):
            gb = g.bounds
            if gb[0] > b[2] or gb[2] < b[0] or gb[1] > b[3] or gb[3] < b[1]:
                continue
            if predicate is None or (predicate == "intersects" and geometry.intersects(g)):
                hits.append(i)
        return np.array(hits, dtype=np.int64)


def nearest_points(a, b):
    """
    ``shapely.ops.nearest_points(area, Point)`` (reference pedestrian/social_force.py:204): the point
    itself when it lies in the closed area, else the closest point on the rings of its polygons
    (GEOS LineSegment::closestPoint: the projection for a projection factor in (0, 1), else the closer
    end point; the first strictly closer segment wins).
    """
    if not isinstance(b, Point):
        raise NotImplementedError(type(b))
    px, py = float(b.x), float(b.y)
    polys = a.geoms if isinstance(a, MultiPolygon) else [a]
    if any(g.point_side(px, py) >= 0 for g in polys):
        return Point(px, py), b
    best, best_d2 = (px, py), float("inf")
    for g in polys:
        for ring in g.rings():
            n = len(ring)
            for k in range(n):
                ax, ay = map(float, ring[k])
                bx, by = map(float, ring[(k + 1) % n])
                dx, dy = bx - ax, by - ay
                len2 = dx * dx + dy * dy
                cx, cy = ax, ay
                if len2 > 0.0:
                    f = ((px - ax) * dx + (py - ay) * dy) / len2
                    if 0.0 < f < 1.0:
                        cx, cy = ax + f * dx, ay + f * dy
                    else:
                        da = (px - ax) * (px - ax) + (py - ay) * (py - ay)
                        db = (px - bx) * (px - bx) + (py - by) * (py - by)
                        if db < da:
                            cx, cy = bx, by
                d2 = (px - cx) * (px - cx) + (py - cy) * (py - cy)
                if d2 < best_d2:
                    best_d2, best = d2, (cx, cy)
    return Point(*best), b


def make_valid(geom):
    return geom


def unary_union(geoms):
    geoms = [g for g in geoms if g is not None]
    polys = []
    for g in geoms:
        if isinstance(g, MultiPolygon):
            polys.extend(g.geoms)
        else:
            polys.append(g)
    return MultiPolygon(polys)
