"""``Scenario`` -- entities + metadata (reference scenario_gym/scenario/scenario.py:20-140)."""
from __future__ import annotations

from copy import copy
from typing import Any, Dict, List, Optional

from .entity import Entity, Pedestrian, Vehicle
from .trajectory import Trajectory


class Scenario:
    """A set of entities with trajectories; the ego is the entity named "ego" or the first."""

    def __init__(self, entities: List[Entity], name: Optional[str] = None, road_network=None,
                 actions: Optional[list] = None, properties: Optional[Dict[Any, Any]] = None):
        self._entities = entities
        self._ref_to_entity: Dict[str, Entity] = {e.ref: e for e in entities}
        self.name = name
        self.road_network = road_network
        self.actions = actions if actions is not None else []
        self.properties = properties if properties is not None else {}

    @property
    def entities(self) -> List[Entity]:
        return self._entities

    @property
    def ego(self) -> Entity:
        ego = self.entity_by_name("ego")
        return ego if ego is not None else self.entities[0]

    @property
    def vehicles(self) -> List[Entity]:
        return [e for e in self.entities if isinstance(e, Vehicle)]

    @property
    def pedestrians(self) -> List[Entity]:
        return [e for e in self.entities if isinstance(e, Pedestrian)]

    @property
    def trajectories(self) -> Dict[str, Trajectory]:
        return {e.ref: e.trajectory for e in self.entities}

    @property
    def length(self) -> float:
        """Largest control-point time over all entities (reference :88-91)."""
        return max(t.max_t for t in self.trajectories.values())

    def entity_by_name(self, e_ref: str) -> Optional[Entity]:
        return self._ref_to_entity.get(e_ref)

    def __copy__(self) -> "Scenario":
        return self.__class__(
            [e.copy() for e in self.entities],
            name=f"Copy of {self.name}" if self.name is not None else None,
            road_network=self.road_network,
            actions=list(self.actions),
            properties=self.properties,
        )

    def copy(self) -> "Scenario":
        return copy(self)
