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

    def surfaces(self) -> Tuple[Surface, Surface, Surface]:
        """(driveable, walkable, impenetrable) in the order the device indexes them."""
        return self.driveable_surface, self.walkable_surface, self.impenetrable_surface
