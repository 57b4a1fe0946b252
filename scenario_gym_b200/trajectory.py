"""
``Trajectory`` -- host-side mirror of the reference's trajectory container
(reference scenario_gym/trajectory.py:12-273).  numpy only; the interpolation kernel
restates scipy's ``interp1d._call_linear`` (``packing.call_linear``) so the values packed for
the device are the ones the reference would compute.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np

from .packing import call_linear

_FIELDS = ("t", "x", "y", "z", "h", "p", "r")


def _resolve_heading(h: np.ndarray) -> np.ndarray:
    """Unwrap headings so there are no large jumps (reference trajectory.py:465-469)."""
    deltas = np.diff(h) % (2 * np.pi)
    deltas = np.where(deltas > np.pi, deltas - 2 * np.pi, deltas)
    return np.hstack([h[0], deltas]).cumsum()


def is_stationary(data: np.ndarray) -> bool:
    """True if every control point has the same pose (reference trajectory.py:472-490)."""
    return len(np.unique(np.where(np.isnan(data[:, 1:]), 0.0, data[:, 1:]), axis=0)) <= 1


class Trajectory:
    """Immutable table of control points ``[t, x, y, z, h, p, r]`` (float64)."""

    _fields = _FIELDS

    def __init__(self, data: np.ndarray, fields: Sequence[str] = _FIELDS):
        fields = tuple(fields)
        if not all(f in fields for f in ("t", "x", "y")):
            raise ValueError("Trajectory cannot be created with t, x and y values.")
        data = np.asarray(data, dtype=np.float64)
        if data.ndim != 2 or data.shape[1] != len(fields):
            raise ValueError(
                f"Invalid shape: {data.shape}. Expected: (N, {len(fields)}). Either pass `fields` to "
                f"specify the columns given or ensure that columns for all of {_FIELDS} are provided."
            )
        perm = [fields.index(f) for f in _FIELDS if f in fields]
        data = data[:, perm]
        data = data[np.unique(data[:, 0], return_index=True)[1]]  # reference :60
        n = data.shape[0]
        cols = []
        for f in _FIELDS:
            d = data[:, perm.index(fields.index(f))] if f in fields else np.zeros(n)
            if f not in fields or np.isfinite(d).sum() != n:
                if f == "h" and n == 1:
                    d = np.zeros(1)
                elif f == "h":  # reference :69-78: heading from a +-1e-2 s finite difference
                    t = cols[0]
                    xy = np.array(cols[1:3]).T
                    diff = call_linear(t, xy, t + 1e-2) - call_linear(t, xy, t - 1e-2)
                    d = _resolve_heading(np.arctan2(diff[:, 1], diff[:, 0]))
                elif f in ("z", "p", "r"):
                    d = np.zeros(n)
                else:
                    raise ValueError(f"Invalid values found for {f}. Values required for xyt.")
            elif f == "h":
                d = _resolve_heading(d)
            cols.append(d)
            setattr(self, f, d)
        self._data = np.array(cols).T.copy()
        self._data.flags.writeable = False

    # ------------------------------------------------------------------ accessors
    @property
    def data(self) -> np.ndarray:
        return self._data

    def __len__(self) -> int:
        return len(self._data)

    def __getitem__(self, idx):
        return self._data[idx]

    @property
    def min_t(self) -> float:
        return self._data[0, 0]

    @property
    def max_t(self) -> float:
        return self._data[-1, 0]

    def is_stationary(self) -> bool:
        return is_stationary(self._data)

    def copy(self) -> "Trajectory":
        return Trajectory(self._data.copy())

    __copy__ = copy

    def to_json(self):
        return self._data.tolist()

    # ------------------------------------------------------------------ queries
    def _interp(self, t: np.ndarray) -> np.ndarray:
        data = self._data
        if data.shape[0] == 1:  # reference :175-177
            data = np.repeat(data, 2, axis=0)
            data[-1, 0] += 1e-3
        return call_linear(data[:, 0], data[:, 1:], np.atleast_1d(t))

    def position_at_t(self, t, extrapolate: Union[bool, Tuple[bool, bool]] = (False, False)
                      ) -> Optional[np.ndarray]:
        """Pose at time t (reference trajectory.py:142-205)."""
        t = np.array(t, dtype=np.float64)
        if isinstance(extrapolate, tuple):
            ext_bck, ext_fwd = extrapolate
            extrapolate = True
        else:
            ext_bck = ext_fwd = extrapolate
        if t.ndim == 0:
            if not extrapolate and (t < self.min_t or t > self.max_t):
                return None
            if t < self.min_t and not ext_bck:
                return self._data[0, 1:]
            if t > self.max_t and not ext_fwd:
                return self._data[-1, 1:]
            return self._interp(t)[0]
        poses = self._interp(t)
        if not ext_bck:
            poses = np.where(t[:, None] < self.min_t, self._data[0, None, 1:], poses)
        if not ext_fwd:
            poses = np.where(t[:, None] > self.max_t, self._data[-1, None, 1:], poses)
        return poses

    def velocity_at_t(self, t, eps: float = 1e-4) -> np.ndarray:
        """Central-difference velocity, zero outside the trajectory (reference :243-273)."""
        t = np.array(t, dtype=np.float64)
        inside = np.logical_and(self.min_t <= t, t <= self.max_t)
        v_in = (self.position_at_t(t + eps / 2, extrapolate=True)
                - self.position_at_t(t - eps / 2, extrapolate=True)) / eps
        v_out = np.zeros(t.shape + (6,))
        if t.ndim >= 1:
            inside = inside.reshape(-1, 1)
        return np.where(inside, v_in, v_out)
