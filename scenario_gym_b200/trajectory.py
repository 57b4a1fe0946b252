"""
``Trajectory`` -- the host-side control-point table of an entity (the container the reference
defines in scenario_gym/trajectory.py; semantics per SURVEY.md section 8 rows a2 / a4 and section 9
item 8).  numpy only.  Interpolation goes through ``packing.call_linear``, the restatement of scipy's
``interp1d._call_linear``, so the values packed for the device are the ones the reference computes.

Semantics kept from the reference (each is observable in rollout results):
  * rows are sorted and de-duplicated by time;
  * missing or non-finite ``z``, ``p``, ``r`` become 0; a missing / non-finite heading is derived
    from the direction of travel (a +-0.01 s central difference of x, y) and every heading column
    is unwrapped, so headings may leave (-pi, pi];
  * a single control point is a "static" trajectory: queries behave as if the point were repeated
    1 ms later;
  * ``position_at_t``: outside the time range the answer is ``None`` (scalar query, no
    extrapolation), the first / last control point (clamped side) or the linear continuation of the
    end segment (extrapolated side);
  * ``velocity_at_t``: a 1e-4 s central difference inside the time range, zero outside.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

from .packing import call_linear

COLUMNS = ("t", "x", "y", "z", "h", "p", "r")
_REQUIRED = ("t", "x", "y")
_ZERO_FILLED = ("z", "p", "r")
_HEADING_PROBE = 1e-2  # seconds either side of a control point when the heading has to be derived
_STATIC_SPAN = 1e-3    # a single control point is treated as two, this far apart


def unwrap_heading(h: np.ndarray) -> np.ndarray:
    """Add multiples of 2 pi so consecutive headings never differ by more than pi."""
    if len(h) < 2:
        return np.array(h, dtype=np.float64)
    step = np.mod(np.diff(h), 2.0 * np.pi)
    step[step > np.pi] -= 2.0 * np.pi
    return np.concatenate(([h[0]], step)).cumsum()


_resolve_heading = unwrap_heading  # the reference's name for it


def is_stationary(data: np.ndarray) -> bool:
    """Do all control points share one pose (NaNs counting as 0)?"""
    poses = np.nan_to_num(np.asarray(data)[:, 1:], nan=0.0, posinf=np.inf, neginf=-np.inf)
    return np.unique(poses, axis=0).shape[0] <= 1


def _travel_heading(t: np.ndarray, xy: np.ndarray) -> np.ndarray:
    """Heading of the direction of travel at every control point (central difference of x, y)."""
    ahead = call_linear(t, xy, t + _HEADING_PROBE)
    behind = call_linear(t, xy, t - _HEADING_PROBE)
    d = ahead - behind
    return np.arctan2(d[:, 1], d[:, 0])


class Trajectory:
    """Immutable ``(K, 7)`` float64 table of control points ``[t, x, y, z, h, p, r]``."""

    _fields = COLUMNS

    def __init__(self, data: np.ndarray, fields: Sequence[str] = COLUMNS):
        given = tuple(fields)
        if any(c not in given for c in _REQUIRED):
            raise ValueError("Trajectory cannot be created with t, x and y values.")
        raw = np.asarray(data, dtype=np.float64)
        if raw.ndim != 2 or raw.shape[1] != len(given):
            raise ValueError(
                f"Invalid shape: {raw.shape}. Expected: (N, {len(given)}). Either pass `fields` to "
                f"specify the columns given or ensure that columns for all of {COLUMNS} are provided."
            )
        # one row per distinct time, in time order
        _, first = np.unique(raw[:, given.index("t")], return_index=True)
        raw = raw[first]
        K = raw.shape[0]
        cols: Dict[str, Optional[np.ndarray]] = {}
        for c in COLUMNS:
            v = raw[:, given.index(c)] if c in given else None
            cols[c] = v if v is not None and bool(np.isfinite(v).all()) else None
        for c in _REQUIRED:
            if cols[c] is None:
                raise ValueError(f"Invalid values found for {c}. Values required for xyt.")
        for c in _ZERO_FILLED:
            if cols[c] is None:
                cols[c] = np.zeros(K)
        if cols["h"] is not None:
            cols["h"] = unwrap_heading(cols["h"])
        elif K == 1:
            cols["h"] = np.zeros(1)
        else:
            cols["h"] = unwrap_heading(_travel_heading(cols["t"], np.column_stack((cols["x"], cols["y"]))))
        table = np.column_stack([cols[c] for c in COLUMNS])
        table.setflags(write=False)
        self._data = table
        for k, c in enumerate(COLUMNS):
            setattr(self, c, table[:, k])

    # ------------------------------------------------------------------ container protocol
    @property
    def data(self) -> np.ndarray:
        return self._data

    def __len__(self) -> int:
        return self._data.shape[0]

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
        return Trajectory(np.array(self._data))

    __copy__ = copy

    def __deepcopy__(self, memo) -> "Trajectory":
        return self.copy()

    def to_json(self):
        return self._data.tolist()

    @classmethod
    def from_json(cls, data) -> "Trajectory":
        return cls(np.array(data, dtype=np.float64))

    # ------------------------------------------------------------------ queries
    def _knots(self) -> Tuple[np.ndarray, np.ndarray]:
        """Times and poses to interpolate over (a static trajectory gets its second knot here)."""
        t, poses = self._data[:, 0], self._data[:, 1:]
        if len(t) == 1:
            t = np.array([t[0], t[0] + _STATIC_SPAN])
            poses = np.vstack((poses, poses))
        return t, poses

    def position_at_t(self, t, extrapolate: Union[bool, Tuple[bool, bool]] = (False, False)
                      ) -> Optional[np.ndarray]:
        """
        Pose(s) at time(s) ``t``.  ``extrapolate`` is one flag for both ends or a
        ``(backwards, forwards)`` pair; a side that is not extrapolated is clamped to its end control
        point -- except that a *scalar* query with the plain flag ``False`` returns ``None`` outside the
        time range (an entity that is not in the scene at that time).
        """
        query = np.asarray(t, dtype=np.float64)
        per_side = isinstance(extrapolate, tuple)
        back, fwd = extrapolate if per_side else (extrapolate, extrapolate)
        before, after = query < self.min_t, query > self.max_t
        if query.ndim == 0:
            if not per_side and not extrapolate and (before or after):
                return None
            if before and not back:
                return self._data[0, 1:]
            if after and not fwd:
                return self._data[-1, 1:]
            return call_linear(*self._knots(), query.reshape(1))[0]
        out = call_linear(*self._knots(), query)
        if not back:
            out[before] = self._data[0, 1:]
        if not fwd:
            out[after] = self._data[-1, 1:]
        return out

    def velocity_at_t(self, t, eps: float = 1e-4) -> np.ndarray:
        """Central-difference velocity of all six pose components; zero outside the time range."""
        query = np.asarray(t, dtype=np.float64)
        half = eps / 2
        rate = (self.position_at_t(query + half, extrapolate=True)
                - self.position_at_t(query - half, extrapolate=True)) / eps
        inside = (self.min_t <= query) & (query <= self.max_t)
        if query.ndim == 0:
            return rate if inside else np.zeros(6)
        rate[~inside] = 0.0
        return rate
