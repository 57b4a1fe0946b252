"""
TEST INFRASTRUCTURE ONLY -- drives ``oracle/_build/libsg_oracle.so`` (the plain-C
restatement of the reference path) through the same ABI as the CUDA engine, with
numpy host buffers.  Imported only by tests/, ``__graft_entry__.smoke`` and
``bench.py``'s cpu_baseline / ``--impl reference`` leg.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from scenario_gym_b200 import abi
from scenario_gym_b200.packing import PackedScene

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "_build", "libsg_oracle.so")

_EVENT_DTYPE = np.dtype(
    [("scenario", "<i4"), ("tick", "<i4"), ("slot", "<i4"), ("_pad", "<i4"), ("t", "<f8")]
)


def build_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "sg_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "sg_b200.h")
    stale = (
        force
        or not os.path.exists(ORACLE_LIB)
        or os.path.getmtime(ORACLE_LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return ORACLE_LIB


_lib: Optional[Dict[str, object]] = None
_cdll = None


def load_oracle() -> Dict[str, object]:
    global _lib, _cdll
    if _lib is None:
        build_oracle()
        _cdll = C.CDLL(ORACLE_LIB)
        _lib = abi.bind(_cdll, "sgo_")
    return _lib


def oracle_cdll():
    load_oracle()
    return _cdll


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def alloc_state(N: int, M: int, event_cap: int = 4096, trace_cap: int = 0,
                coll_matrix: bool = True) -> Dict[str, np.ndarray]:
    """Allocate every SgState array as numpy (host)."""
    W = (M + 31) // 32
    dims = {"N": N, "M": M, "NM": N * M, "W": W, "E": max(event_cap, 1), "T": max(trace_cap, 1)}
    out = {}
    for name, dtype, shape in abi.STATE_FIELDS:
        shp = tuple(dims[s] if s in dims else int(s) for s in shape)
        if name == "coll_mask" and not coll_matrix:
            shp = (1,)
        out[name] = np.zeros(shp, _EVENT_DTYPE if dtype == "event" else np.dtype(dtype))
    out["_event_cap"] = event_cap
    out["_trace_cap"] = trace_cap
    return out


def scene_struct(scene: PackedScene) -> abi.SgScene:
    s = abi.SgScene()
    s.n_scenarios, s.n_slots = scene.N, scene.M
    s.n_traj_rows = scene.traj_rows.shape[0]
    s.n_union_rows = scene.union_t.shape[0]
    s.n_route_pts = scene.route_xy.shape[0]
    s.n_networks = scene.n_networks
    s.n_rn_polys = len(scene.rn_edge_off) - 1
    s.n_rn_edges = scene.rn_edges.shape[0]
    s.kind_mask = scene.kind_mask()
    s.scene_flags = scene.scene_flags()
    for k, a in scene.arrays().items():
        assert a.flags["C_CONTIGUOUS"], k
        setattr(s, k, a.ctypes.data)
    return s


def state_struct(state: Dict[str, np.ndarray]) -> abi.SgState:
    s = abi.SgState()
    for name, _, _ in abi.STATE_FIELDS:
        setattr(s, name, state[name].ctypes.data)
    s.event_cap = state["_event_cap"]
    s.trace_cap = state["_trace_cap"]
    return s


class OracleEngine:
    """reset()/rollout() on the CPU oracle; mirrors scenario_gym_b200.engine.Engine."""

    def __init__(self, scene: PackedScene, params: abi.SgParams, event_cap: int = 4096,
                 trace_cap: int = 0):
        self.lib = load_oracle()
        self.scene = scene
        self.params = params
        self.state = alloc_state(scene.N, scene.M, event_cap, trace_cap)
        self._sc = scene_struct(scene)
        self._st = state_struct(self.state)

    def reset(self):
        rc = self.lib["reset"](C.byref(self._sc), C.byref(self.params), C.byref(self._st), 0, None)
        if rc:
            raise RuntimeError(self.lib["last_error"]().decode())

    def rollout(self, n_ticks: int = -1, actions: Optional[np.ndarray] = None,
                host_pose: Optional[np.ndarray] = None, host_present: Optional[np.ndarray] = None,
                step_done: bool = False, tick0: int = 0):
        inp = abi.SgInputs()
        inp.step_done = int(step_done)
        from scenario_gym_b200.action_rng import ActionRng

        if isinstance(actions, ActionRng):  # the oracle's own restatement of numpy's PCG64 stream
            assert actions.nm == self.scene.N * self.scene.M
            inp.use_rng, inp.rng_tick0 = 1, int(tick0)
            inp.rng = actions.struct()
            inp.n_action_ticks = max(actions.n_ticks - tick0, 0)
        elif actions is not None:
            if getattr(actions, "dtype", None) == np.float32:
                actions = np.ascontiguousarray(actions[tick0:])
                inp.actions_f32 = actions.ctypes.data
            else:
                actions = np.ascontiguousarray(actions[tick0:], np.float64)
                inp.actions = actions.ctypes.data
            assert actions.shape[1:] == (2, self.scene.N * self.scene.M), actions.shape
            inp.n_action_ticks = actions.shape[0]
        if host_pose is not None:
            inp.host_pose = host_pose.ctypes.data
            inp.host_present = host_present.ctypes.data
        rc = self.lib["rollout"](
            C.byref(self._sc), C.byref(self.params), C.byref(self._st), C.byref(inp), n_ticks, 0, None
        )
        if rc:
            raise RuntimeError(self.lib["last_error"]().decode())

    def fill_actions(self, rng, tick0: int = 0, n_ticks: Optional[int] = None) -> np.ndarray:
        n_ticks = rng.n_ticks - tick0 if n_ticks is None else n_ticks
        out = np.empty((n_ticks, 2, rng.nm))
        r = rng.struct()
        rc = self.lib["fill_random_actions"](C.byref(r), int(tick0), int(n_ticks), int(rng.nm),
                                             out.ctypes.data, 0, None)
        if rc:
            raise RuntimeError(self.lib["last_error"]().decode())
        return out

    def future_collisions(self, t=None, horizon: float = 5.0, n_samples: int = 10, slot=None) -> np.ndarray:
        tt = np.ascontiguousarray(self.state["t"] if t is None else t, np.float64)
        sl = None if slot is None else np.ascontiguousarray(slot, np.int32)
        out = np.zeros(self.scene.N, np.uint8)
        rc = self.lib["future_collisions"](C.byref(self._sc), tt.ctypes.data, _ptr(sl), float(horizon),
                                           int(n_samples), out.ctypes.data, 0, None)
        if rc:
            raise RuntimeError(self.lib["last_error"]().decode())
        return out.astype(bool)

    def entities_in_radius(self, x, y, r) -> np.ndarray:
        N, M = self.scene.N, self.scene.M
        a = [np.ascontiguousarray(np.broadcast_to(np.asarray(v, np.float64), (N,))) for v in (x, y, r)]
        out = np.zeros(N * M, np.uint8)
        rc = self.lib["entities_in_radius"](C.byref(self._st), N, M, a[0].ctypes.data, a[1].ctypes.data,
                                            a[2].ctypes.data, out.ctypes.data, 0, None)
        if rc:
            raise RuntimeError(self.lib["last_error"]().decode())
        return out.astype(bool).reshape(N, M)

    def events(self) -> np.ndarray:
        n = min(int(self.state["event_count"][0]), self.state["_event_cap"])
        ev = self.state["events"][:n]
        order = np.lexsort((ev["slot"], ev["tick"], ev["scenario"]))
        return ev[order]


def _get(self, name: str) -> np.ndarray:
    """Host copy of a state array (same accessor as the CUDA Engine)."""
    return self.state[name]


OracleEngine.get = _get
