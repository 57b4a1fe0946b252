"""
Entities, catalog entries and bounding boxes -- host-side mirror of the reference's data
model (reference scenario_gym/entity/base.py, catalog_entry.py:83-247, entity/vehicle.py,
entity/pedestrian.py, entity/misc.py).  Only what the rollout path reads is kept.
"""
from __future__ import annotations

from copy import copy
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

from . import abi
from .trajectory import Trajectory


@dataclass
class BoundingBox:
    """Box given by width, length and the centre offset from the reference point."""

    width: float
    length: float
    center_x: float
    center_y: float

    def to_dict(self):
        return {"width": self.width, "length": self.length, "center_x": self.center_x,
                "center_y": self.center_y}

    @classmethod
    def from_dict(cls, d):
        return cls(d["width"], d["length"], d["center_x"], d["center_y"])


@dataclass
class CatalogEntry:
    """Catalog information of an entity (reference catalog_entry.py:140-247)."""

    catalog: Optional[Any]
    catalog_entry: str
    catalog_category: Optional[str]
    catalog_type: str
    bounding_box: BoundingBox
    properties: Dict[str, Any] = field(default_factory=dict)
    files: List[str] = field(default_factory=list)

    def to_dict(self) -> Dict[str, Any]:
        cat = self.catalog
        if cat is not None and not isinstance(cat, dict):
            cat = {"name": getattr(cat, "name", str(cat)), "group_name": getattr(cat, "group_name", None)}
        return {"catalog": cat, "catalog_entry": self.catalog_entry, "catalog_category": self.catalog_category,
                "catalog_type": self.catalog_type, "bounding_box": self.bounding_box.to_dict(),
                "properties": self.properties, "files": self.files}

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "CatalogEntry":
        return cls(d.get("catalog"), d["catalog_entry"], d["catalog_category"], d["catalog_type"],
                   BoundingBox.from_dict(d["bounding_box"]), d.get("properties", {}), d.get("files", []))


class Entity:
    """An entity: a catalog entry plus a trajectory (reference entity/base.py:15-183)."""

    def __init__(self, catalog_entry: CatalogEntry, trajectory: Optional[Trajectory] = None,
                 ref: Optional[str] = None):
        self.ref = ref
        self.catalog_entry = catalog_entry
        self._trajectory = trajectory

    @property
    def trajectory(self) -> Trajectory:
        return self._trajectory

    @trajectory.setter
    def trajectory(self, trajectory: Trajectory) -> None:
        self._trajectory = trajectory

    @property
    def bounding_box(self) -> BoundingBox:
        return self.catalog_entry.bounding_box

    @property
    def type(self) -> Optional[str]:
        return self.catalog_entry.catalog_type.replace("Catalogs", "")

    def is_static(self) -> bool:
        return self.trajectory.data.shape[0] == 1

    def __copy__(self) -> "Entity":
        return self.__class__(
            self.catalog_entry,
            trajectory=None if self.trajectory is None else self.trajectory.copy(),
            ref=self.ref,
        )

    def copy(self) -> "Entity":
        return copy(self)

    def to_dict(self) -> Dict[str, Any]:
        """reference entity/base.py:158-165"""
        return {"ref": self.ref, "trajectory": self.trajectory.to_json(),
                "catalog_entry": self.catalog_entry.to_dict(), "entity_class": self.__class__.__name__}

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "Entity":
        return cls(CatalogEntry.from_dict(data["catalog_entry"]),
                   trajectory=Trajectory(np.array(data["trajectory"])), ref=data.get("ref"))

    def get_bounding_box_points(self, pose) -> np.ndarray:
        """Corners RL, FL, FR, RR in the global frame (reference entity/base.py:100-138)."""
        pose = np.asarray(pose, dtype=np.float64)
        ref_xy, h = pose[..., :2], pose[..., 3 if pose.shape[-1] > 3 else 2]
        n = h.ndim
        R = np.array([[np.cos(h), np.sin(h)], [-np.sin(h), np.cos(h)]]).transpose(
            *(tuple(i + 2 for i in range(n)) + (0, 1)))
        bb = self.bounding_box
        points = np.array([
            [bb.center_x - 0.5 * bb.length, bb.center_y + 0.5 * bb.width],
            [bb.center_x + 0.5 * bb.length, bb.center_y + 0.5 * bb.width],
            [bb.center_x + 0.5 * bb.length, bb.center_y - 0.5 * bb.width],
            [bb.center_x - 0.5 * bb.length, bb.center_y - 0.5 * bb.width],
        ])
        return ref_xy[..., None, :] + np.einsum("ij,...jk->...ik", points, R)

    # engine-facing classification
    def etype(self) -> int:
        if isinstance(self, Vehicle):
            return abi.ETYPE_VEHICLE
        if isinstance(self, Pedestrian) or self.type == "Pedestrian":
            return abi.ETYPE_PEDESTRIAN
        return abi.ETYPE_MISC


class Vehicle(Entity):
    """Entity loaded from a ``Vehicle`` catalog element."""


class Pedestrian(Entity):
    """Entity loaded from a ``Pedestrian`` catalog element."""


class MiscObject(Entity):
    """Entity loaded from a ``MiscObject`` catalog element."""


ENTITY_CLASS_BY_TAG = {"Vehicle": Vehicle, "Pedestrian": Pedestrian, "MiscObject": MiscObject}
