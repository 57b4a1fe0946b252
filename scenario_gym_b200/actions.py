"""
Scenario actions (reference scenario_gym/scenario/actions.py:12-170): events attached to a scenario
that fire once their trigger condition holds and write into ``state.entity_state``.  They are host
objects; ``State.update_actions`` applies them after every tick -- or, after a fused rollout, for the
whole sequence of tick times at once (the tick times are re-derived by the same repeated addition
the device performs, so ``action_apply_times`` are the times the reference records).
"""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Dict, Optional

import numpy as np


class ScenarioAction:
    """Base class: subclasses implement ``_apply(state, entity)`` and ``trigger_condition(state)``."""

    def __init__(self, action_class: str, entity_ref: str, action_variables: Dict[str, Any]):
        self.action_class = action_class
        self.entity_ref = entity_ref
        self.action_variables = action_variables

    def apply(self, state, entity) -> None:
        self._apply(state, entity)

    def _apply(self, state, entity) -> None:
        raise NotImplementedError

    def trigger_condition(self, state) -> bool:
        raise NotImplementedError

    def copy(self):
        return deepcopy(self)

    def translate(self, x: np.ndarray, inplace: bool = False):
        return self if inplace else self.copy()

    def to_dict(self) -> Dict[str, Any]:
        return {"action_class": self.action_class, "entity_ref": self.entity_ref,
                "action_variables": self.action_variables}

    @classmethod
    def from_dict(cls, data: Dict[str, Any]):
        return cls(data["action_class"], data["entity_ref"], data["action_variables"])


class FixedTAction(ScenarioAction):
    """Fires at the first tick whose time is at or after ``t``."""

    def __init__(self, t: float, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.t = t

    def trigger_condition(self, state) -> bool:
        return state.t >= self.t

    def translate(self, x: np.ndarray, inplace: bool = False):
        act = self if inplace else self.copy()
        act.t += x[0]
        return act

    def to_dict(self) -> Dict[str, Any]:
        data = super().to_dict()
        data["t"] = self.t
        return data

    @classmethod
    def from_dict(cls, data: Dict[str, Any]):
        return cls(data["t"], data["action_class"], data["entity_ref"], data["action_variables"])


class UserDefinedAction(FixedTAction):
    """An OpenSCENARIO UserDefinedAction: carried along, applies nothing."""

    def _apply(self, state, entity) -> None:
        pass


class UpdateStateVariableAction(FixedTAction):
    """Writes its variables into ``state.entity_state[entity]`` once ``state.t`` has passed ``t``."""

    def _apply(self, state, entity) -> None:
        if entity is not None:
            if state.entity_state[entity] is None:
                state.entity_state[entity] = {}
            for k, v in self.action_variables.items():
                state.entity_state[entity][k] = v

    def trigger_condition(self, state) -> bool:
        return state.t > self.t
