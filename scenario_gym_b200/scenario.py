"""``Scenario`` -- entities + metadata (reference scenario_gym/scenario/scenario.py:20-140)."""
from __future__ import annotations

import json
import os
from copy import copy
from typing import Any, Dict, List, Optional, Tuple, Type

from .actions import ScenarioAction, UpdateStateVariableAction
from .entity import Entity, MiscObject, Pedestrian, Vehicle
from .road_network import RoadNetwork
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

    def add_action(self, action: ScenarioAction, inplace: bool = False) -> "Scenario":
        scenario = self if inplace else self.copy()
        scenario.actions.append(action)
        return scenario

    def translate(self, x, inplace: bool = False) -> "Scenario":
        """Shift every trajectory (and action time) by x = [t, x, y, z, h, p, r] (reference scenario.py:166-177)."""
        import numpy as np

        scenario = self if inplace else self.copy()
        x = np.asarray(x, dtype=np.float64)
        for e in scenario.entities:
            e.trajectory = Trajectory(e.trajectory.data + x[None, :])
        scenario.actions = [a.translate(x, inplace=inplace) for a in scenario.actions]
        return scenario

    def reset_start(self, entity: Optional[Entity] = None) -> "Scenario":
        import numpy as np

        entity = self.ego if entity is None else entity
        return self.translate(np.array([-entity.trajectory.min_t, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]))

    # ------------------------------------------------------------------ json (reference scenario.py:186-319)
    def to_dict(self, road_network_path: Optional[str] = "../Road_Networks") -> Dict[str, Any]:
        rn = self.road_network
        if rn is None:
            road_network = None
        else:
            path = getattr(rn, "path", None)
            if path is None and road_network_path is not None:
                path = road_network_path if os.path.isfile(road_network_path) else \
                    os.path.join(road_network_path, f"{rn.name}.json")
            road_network = {"path": path, "name": rn.name}
        return {"entities": [e.to_dict() for e in self.entities], "name": self.name,
                "actions": [a.to_dict() for a in self.actions], "road_network": road_network,
                "properties": self.properties}

    @classmethod
    def from_dict(cls, data: Dict[str, Any], e_classes: Tuple[Type[Entity], ...] = (Vehicle, Pedestrian, MiscObject, Entity),
                  a_classes: Tuple[Type[ScenarioAction], ...] = (UpdateStateVariableAction,)) -> "Scenario":
        entities = []
        for e_data in data["entities"]:
            Ent = next((E for E in e_classes if E.__name__ == e_data.get("entity_class")), Entity)
            entities.append(Ent.from_dict(e_data))
        road_network = data.get("road_network")
        if road_network is not None:
            path = road_network.get("path") if isinstance(road_network, dict) else None
            if path is not None:
                if os.path.exists(path):
                    road_network = RoadNetwork.create_from_file(path)
                elif road_network.get("name") is not None:
                    road_network = RoadNetwork(name=road_network["name"])
                else:
                    road_network = None
            else:
                road_network = RoadNetwork.create_from_dict(road_network)
        actions = []
        for a_data in data.get("actions", ()):
            Act = next((A for A in a_classes if A.__name__ == a_data.get("action_class")), a_classes[-1])
            actions.append(Act.from_dict(a_data))
        return cls(entities, name=data.get("name"), road_network=road_network, actions=actions,
                   properties=data.get("properties", {}))

    @classmethod
    def from_json(cls, path: str, road_network_dir: Optional[str] = None, **kwargs) -> "Scenario":
        with open(path) as f:
            data = json.load(f)
        rn = data.get("road_network")
        if isinstance(rn, dict) and rn.get("path") is not None and not os.path.isabs(rn["path"]):
            base = os.path.dirname(os.path.abspath(path))
            if road_network_dir is not None:
                base = road_network_dir if os.path.isabs(road_network_dir) else os.path.join(base, road_network_dir)
                rn["path"] = os.path.join(base, os.path.basename(rn["path"]))
            else:
                rn["path"] = os.path.join(base, rn["path"])
        if data.get("name") is None:
            data["name"] = os.path.splitext(os.path.basename(path))[0]
        return cls.from_dict(data, **kwargs)

    def to_json(self, path: str, road_network_path: Optional[str] = "../Road_Networks") -> None:
        with open(path, "w") as f:
            json.dump(self.to_dict(road_network_path=road_network_path), f)

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
