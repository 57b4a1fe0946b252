"""
``State`` -- the read surface plugins rely on (reference scenario_gym/state/state.py), backed
by the device buffers of one scenario.  It is materialised lazily (one small device->host copy
per tick) and only when a host-side plugin, metric or terminal condition needs it.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .entity import Entity


class State:
    """Global state of one scenario: time, terminal flag, poses and velocities of the entities."""

    def __init__(self, gym, n: int, scenario, scenario_path: Optional[str] = None):
        self._gym = gym
        self._n = n
        self._scenario = scenario
        self.scenario_path = scenario_path
        self.persist = gym.persist
        self.agents: Dict[Entity, Any] = {}
        self.state_callbacks = gym.state_callbacks
        self.entity_state: Dict[Entity, Any] = dict.fromkeys(scenario.entities)
        self.unapplied_actions: list = []
        self.action_apply_times: dict = {}
        self.next_t: Optional[float] = None
        self.last_keystroke = None
        self._recorded: Dict[Entity, List[Tuple[float, np.ndarray]]] = {e: [] for e in scenario.entities}
        self._snap: Optional[dict] = None
        self._collisions = None
        self._reset_actions()

    # ------------------------------------------------------------------ scenario actions
    def _reset_actions(self) -> None:
        """reference state.py:145-163 (_reset_data): nothing applied, entity states cleared."""
        actions = list(getattr(self._scenario, "actions", None) or [])
        self.unapplied_actions = actions.copy()
        self.action_apply_times = {a: float("nan") for a in actions}
        self.entity_state = dict.fromkeys(self._scenario.entities)

    def update_actions(self) -> None:
        """Apply the actions whose trigger condition holds now (reference state.py:241-250)."""
        if not self.unapplied_actions:
            return
        still = []
        for act in self.unapplied_actions:
            if act.trigger_condition(self):
                self.apply_action(act)
                self.action_apply_times[act] = self.t
            else:
                still.append(act)
        self.unapplied_actions = still

    def apply_action(self, action) -> None:
        import warnings

        entity = self._scenario.entity_by_name(action.entity_ref)
        if entity is None:
            warnings.warn(f"No entity with name {action.entity_ref} was found for action "
                          f"{action.__class__.__name__}.")
        else:
            action.apply(self, entity)

    def _replay_actions(self, times) -> None:
        """After a fused rollout: visit the tick times one by one so every action fires at its tick."""
        if not self.unapplied_actions:
            return
        saved = self._snap
        for t in times:
            if not self.unapplied_actions:
                break
            self._snap = {"t": float(t), "prev_t": float(t), "done": False, "poses": {}, "velocities": {},
                          "distances": {}}
            self.update_actions()
        self._snap = saved

    # ------------------------------------------------------------------ device sync
    def _invalidate(self) -> None:
        self._snap = None
        self._collisions = None

    def _data(self) -> dict:
        if self._snap is None:
            self._snap = self._gym._materialise(self._n)
        return self._snap

    def _record(self) -> None:
        """Append the current poses to the host-side trace (State._recorded_poses)."""
        d = self._data()
        for e, pose in d["poses"].items():
            self._recorded[e].append((d["t"], pose))

    # ------------------------------------------------------------------ reference surface
    @property
    def scenario(self):
        return self._scenario

    @property
    def t(self) -> float:
        return self._data()["t"]

    @property
    def prev_t(self) -> float:
        return self._data()["prev_t"]

    @property
    def dt(self) -> float:
        d = self._data()
        return d["t"] - d["prev_t"]

    @property
    def is_done(self) -> bool:
        return self._data()["done"] or self._gym._host_done[self._n]

    @property
    def poses(self) -> Dict[Entity, np.ndarray]:
        return self._data()["poses"]

    @property
    def velocities(self) -> Dict[Entity, np.ndarray]:
        return self._data()["velocities"]

    @property
    def distances(self) -> Dict[Entity, float]:
        return self._data()["distances"]

    def collisions(self) -> Dict[Entity, List[Entity]]:
        """Entities whose boxes intersect at the current time (reference state.py:306-310)."""
        if self._collisions is None:
            self._collisions = self._gym._collisions(self._n)
        return self._collisions

    def get_callback(self, Callback):
        for cb in self.state_callbacks:
            if isinstance(cb, Callback):
                return cb
        return None

    def recorded_poses(self, entity: Optional[Entity] = None):
        def table(rows):
            if not rows:
                return np.empty((0, 7))
            ts, poses = map(np.array, zip(*rows))
            return np.concatenate([ts[:, None], poses], axis=1)

        dev = self._gym._device_trace(self._n)
        if dev is not None:  # fused rollout with ScenarioGym(record=True)
            return dev[entity] if entity is not None else dev
        if entity is not None:
            return table(self._recorded.get(entity))
        return {e: table(r) for e, r in self._recorded.items()}

    def to_scenario(self, name: Optional[str] = None):
        """Scenario built from the recorded poses (reference state/state.py:374-394)."""
        from copy import deepcopy

        from .scenario import Scenario
        from .trajectory import Trajectory, is_stationary

        if name is None:
            name = f"Simulation of {self.scenario.name}" if self.scenario.name is None else None
        entities = []
        for entity, poses in self.recorded_poses().items():
            new_entity = deepcopy(entity)
            if is_stationary(poses):
                poses = poses[None, 0]
            new_entity.trajectory = Trajectory(poses)
            entities.append(new_entity)
        return Scenario(entities, name=name, road_network=self.scenario.road_network,
                        actions=self.scenario.actions)

    def get_entity_data(self, entity: Entity):
        return (self.t, self.next_t, self.poses.get(entity), self.velocities.get(entity),
                self.distances.get(entity), self.recorded_poses(entity=entity),
                self.entity_state.get(entity))

    def get_entity_box_points(self, e: Entity) -> np.ndarray:
        return e.get_bounding_box_points(self.poses[e])

    def get_road_info_at_entity(self, e: Entity):
        """Names and geometries of the road network at the entity's position (reference state.py:330-338)."""
        rn = self.scenario.road_network
        if not rn:
            return [], []
        return rn.get_geometries_at_point(*self.poses[e][:2])

    def get_entities_in_radius(self, x: float, y: float, r: float) -> List[Entity]:
        """
        Entities whose position lies strictly inside ``Point(x, y).buffer(r)`` -- the 64-gon GEOS
        builds, not the circle (reference state.py:352-372).  Decided on the device for the whole
        batch (``sg_entities_in_radius``) with the sensor's own predicate.
        """
        mask = self._gym._entities_in_radius(self._n, float(x), float(y), float(r))
        ents = self._gym._entity_of[self._n]
        return [e for s, e in enumerate(ents) if mask[s]]

    def get_entities_in_area(self, area) -> List[Entity]:
        """
        Entities whose position lies strictly inside ``area`` (reference state.py:340-350): a
        ``road_network.PolygonArea`` / ``Surface``, or a sequence of (x, y) vertices.
        """
        from .road_network import PolygonArea, Surface

        if not isinstance(area, (PolygonArea, Surface)):
            area = PolygonArea(area)
        return [e for e, pose in self.poses.items() if area.contains(pose[0], pose[1])]
