"""
``RoadNetwork`` -- the part of the reference's road network (scenario_gym/road_network/) the rollout
path touches: the driveable / walkable / impenetrable *surfaces* behind the ``ego_off_road``
terminal condition (state/state.py:401-407) and the social-force boundary forces
(pedestrian/social_force.py:83-104, 190-211).

The reference builds the surfaces with ``shapely.ops.unary_union`` over the member geometries'
boundaries (road_network/road_network.py:306-328).  Here a surface stays a polygon *soup*
(``Surface``): containment is "strictly inside one member", the nearest point is taken over the
members' rings -- what the union answers, except for points exactly on an edge two members share.
The same soup is packed for the device (``packing.pack_road_networks``).

Everything else the reference's class offers (lane graphs, elevation, OpenDRIVE import,
rasterisation) is outside the per-tick path and not mirrored.
"""
from __future__ import annotations

import json
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


def _ring(points) -> np.ndarray:
    pts = np.asarray([(float(p["x"]), float(p["y"])) if isinstance(p, dict) else (float(p[0]), float(p[1]))
                      for p in points], np.float64).reshape(-1, 2)
    if len(pts) > 1 and (pts[0] == pts[-1]).all():
        pts = pts[:-1]
    return pts


def _orient(ax, ay, bx, by, cx, cy) -> int:
    """Exact sign of the orientation determinant (floating-point filter, rational arithmetic behind it)."""
    l, r = (ax - cx) * (by - cy), (ay - cy) * (bx - cx)
    det = l - r
    if abs(det) > 3.3306690738754716e-16 * (abs(l) + abs(r)):
        return 1 if det > 0 else -1
    from fractions import Fraction as F

    d = (F(ax) - F(cx)) * (F(by) - F(cy)) - (F(ay) - F(cy)) * (F(bx) - F(cx))
    return (d > 0) - (d < 0)


def _ring_side(pts: np.ndarray, px: float, py: float) -> int:
    """+1 strictly inside the ring, 0 on it, -1 outside (crossing parity, exact)."""
    inside = False
    n = len(pts)
    for k in range(n):
        ax, ay = float(pts[k, 0]), float(pts[k, 1])
        bx, by = float(pts[(k + 1) % n, 0]), float(pts[(k + 1) % n, 1])
        straddles = (ay > py) != (by > py)
        in_box = min(ax, bx) <= px <= max(ax, bx) and min(ay, by) <= py <= max(ay, by)
        if not straddles and not in_box:
            continue
        o = _orient(ax, ay, bx, by, px, py)
        if o == 0 and in_box:
            return 0
        if straddles and ((o > 0) == (by > ay)):
            inside = not inside
    return 1 if inside else -1


class PolygonArea:
    """A polygon with optional holes, as rings of (x, y) vertices (open rings: no repeated end point)."""

    def __init__(self, exterior, interiors: Iterable = ()):
        self.exterior = _ring(exterior)
        self.interiors = [_ring(r) for r in interiors]

    @classmethod
    def from_json(cls, boundary) -> "PolygonArea":
        """road_network/utils.py:6-26: a list of {x, y} or {"exterior": [...], "interiors": [[...]]}."""
        if isinstance(boundary, dict):
            return cls(boundary["exterior"], boundary.get("interiors", ()))
        return cls(boundary)

    def rings(self) -> List[np.ndarray]:
        return [self.exterior] + self.interiors

    def edges(self) -> np.ndarray:
        """(E, 4) rows x0, y0, x1, y1 over every ring."""
        out = []
        for r in self.rings():
            if len(r) >= 2:
                out.append(np.concatenate([r, np.roll(r, -1, axis=0)], axis=1))
        return np.concatenate(out, axis=0) if out else np.zeros((0, 4))

    def point_side(self, x: float, y: float) -> int:
        """+1 interior, 0 boundary, -1 exterior (the same crossing-parity rule the device applies)."""
        x, y = float(x), float(y)
        side = _ring_side(self.exterior, x, y)
        if side <= 0:
            return side
        for h in self.interiors:
            hs = _ring_side(h, x, y)
            if hs == 0:
                return 0
            if hs > 0:
                return -1
        return 1

    def contains(self, x: float, y: float) -> bool:
        return self.point_side(x, y) > 0

    @property
    def area(self) -> float:
        def shoelace(r):
            return 0.5 * abs(float(np.dot(r[:, 0], np.roll(r[:, 1], -1)) - np.dot(np.roll(r[:, 0], -1), r[:, 1])))

        return shoelace(self.exterior) - sum(shoelace(r) for r in self.interiors)


class Surface:
    """Union of polygons, kept as a soup (see the module docstring)."""

    def __init__(self, polygons: Sequence[PolygonArea] = ()):
        self.polygons = list(polygons)

    @property
    def area(self) -> float:
        return float(sum(p.area for p in self.polygons))

    def contains(self, x: float, y: float) -> bool:
        """Strictly inside one member polygon."""
        return any(p.contains(x, y) for p in self.polygons)

    def __len__(self) -> int:
        return len(self.polygons)


class RoadObject:
    def __init__(self, id: str):
        self.id = id

    def __eq__(self, other) -> bool:
        if isinstance(other, str):
            return self.id == other
        return hasattr(other, "id") and other.id == self.id

    def __hash__(self) -> int:
        return hash(self.id)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(id={self.id})"


class RoadGeometry(RoadObject):
    """A geometry with a boundary polygon; the class flags say which surfaces it joins (base.py:52-68)."""

    driveable = True
    walkable = True
    impenetrable = False

    def __init__(self, id: str, boundary, **_ignored):
        super().__init__(id)
        self.boundary = boundary if isinstance(boundary, PolygonArea) else PolygonArea.from_json(boundary)

    @classmethod
    def from_dict(cls, data: Dict[str, Any]):
        return cls(data["Id" if "Id" in data else "id"], PolygonArea.from_json(data["Boundary"]))


class Lane(RoadGeometry):  # objects.py:53
    walkable = False


class Road(RoadGeometry):  # objects.py:117
    walkable = False

    def __init__(self, id: str, boundary, lanes: Sequence[Lane] = (), **kw):
        super().__init__(id, boundary)
        self.lanes = list(lanes)

    @classmethod
    def from_dict(cls, data):
        return cls(data["Id" if "Id" in data else "id"], PolygonArea.from_json(data["Boundary"]),
                   lanes=[Lane.from_dict(l) for l in data.get("Lanes", [])])


class Intersection(Road):  # objects.py:152-153
    driveable = True
    walkable = False


class Pavement(RoadGeometry):  # objects.py:193
    driveable = False


class Crossing(RoadGeometry):  # objects.py:203
    driveable = False


class Building(RoadGeometry):  # objects.py:240-241
    driveable = False
    impenetrable = True


class RoadNetwork:
    """Collection of road geometries (reference road_network/road_network.py:29-60, 306-328)."""

    _default_object_names = {"roads": Road, "intersections": Intersection, "lanes": Lane,
                             "pavements": Pavement, "crossings": Crossing, "buildings": Building}

    def __init__(self, name: Optional[str] = None, path: Optional[str] = None, **objects):
        self.name, self.path = name, path
        self.object_names = dict(self._default_object_names)
        for k in self._default_object_names:
            setattr(self, k, list(objects.pop(k, []) or []))
        for k, v in objects.items():  # custom layers of RoadObject subclasses
            v = list(v or [])
            setattr(self, k, v)
            if v:
                self.object_names[k] = type(v[0])
        # the lanes layer = the lanes of the roads and intersections plus the ones passed (road_network.py:280-286)
        nested = [l for r in list(self.roads) + list(self.intersections) for l in getattr(r, "lanes", [])]
        self.lanes = list(dict.fromkeys(nested + list(self.lanes)))
        self._surfaces: Dict[str, Surface] = {}

    # ------------------------------------------------------------------ construction
    @classmethod
    def create_from_json(cls, filepath: str) -> "RoadNetwork":
        with open(filepath) as f:
            data = json.load(f)
        return cls.from_dict(data, path=filepath)

    @classmethod
    def create_from_file(cls, filepath: str) -> "RoadNetwork":
        if str(filepath).endswith(".json"):
            return cls.create_from_json(filepath)
        raise NotImplementedError("only JSON road networks are read (OpenDRIVE import is outside the rollout path)")

    @classmethod
    def create_from_dict(cls, data: Dict[str, Any], **kwargs) -> "RoadNetwork":
        return cls.from_dict(data, **kwargs)

    @classmethod
    def from_dict(cls, data: Dict[str, Any], path: Optional[str] = None, name: Optional[str] = None) -> "RoadNetwork":
        layers = {}
        for key, klass in cls._default_object_names.items():
            raw = data.get(key.capitalize(), data.get(key, []))
            layers[key] = [klass.from_dict(d) for d in raw or []]
        return cls(name=name or data.get("name"), path=path, **layers)

    # ------------------------------------------------------------------ surfaces
    @property
    def road_network_geometries(self) -> List[RoadGeometry]:
        out: List[RoadGeometry] = []
        for name, klass in self.object_names.items():
            if issubclass(klass, RoadGeometry):
                out.extend(getattr(self, name))
        return out

    def _surface(self, flag: str) -> Surface:
        if flag not in self._surfaces:
            self._surfaces[flag] = Surface([g.boundary for g in self.road_network_geometries if getattr(g, flag)])
        return self._surfaces[flag]

    @property
    def driveable_surface(self) -> Surface:
        return self._surface("driveable")

    @property
    def walkable_surface(self) -> Surface:
        return self._surface("walkable")

    @property
    def impenetrable_surface(self) -> Surface:
        return self._surface("impenetrable")

    def get_geometries_at_point(self, x: float, y: float) -> Tuple[List[str], List[RoadGeometry]]:
        """Class names and objects of the geometries strictly containing (x, y) (road_network.py:375-401)."""
        names, geoms = [], []
        for g in self.road_network_geometries:
            if g.boundary.contains(x, y):
                names.append(g.__class__.__name__)
                geoms.append(g)
        return names, geoms

    def surfaces(self) -> Tuple[Surface, Surface, Surface]:
        """(driveable, walkable, impenetrable) in the order the device indexes them."""
        return self.driveable_surface, self.walkable_surface, self.impenetrable_surface
