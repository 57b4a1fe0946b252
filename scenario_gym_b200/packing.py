"""
Scene packing: lower scenarios + agent assignment to the flat structure-of-arrays
description consumed by the engine (``SgScene`` in include/sg_b200.h).

Host-side logic only (numpy).  The one piece of reference arithmetic that lives
here is the construction of the ``BatchReplayEntity`` union-knot table
(reference entity/batch.py:55-128), which the device then interpolates per tick.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import abi


def call_linear(x: np.ndarray, y: np.ndarray, x_new: np.ndarray) -> np.ndarray:
    """
    Linear interpolation with the operation order of scipy 1.18 ``interp1d._call_linear``
    (scipy/interpolate/_interpolate.py:491-518), as used by the reference at
    trajectory.py:178-184 and entity/batch.py:99-127.

    x: (K,), y: (K, m), x_new: (n,) -> (n, m).  Extrapolates linearly outside [x0, xK-1].
    """
    idx = np.searchsorted(x, x_new).clip(1, len(x) - 1).astype(int)
    lo, hi = idx - 1, idx
    x_lo, x_hi = x[lo], x[hi]
    return ((x_new - x_lo) / (x_hi - x_lo))[:, None] * y[hi] + (
        (x_hi - x_new) / (x_hi - x_lo)
    )[:, None] * y[lo]


def call_linear_clamped(x: np.ndarray, y: np.ndarray, x_new: np.ndarray) -> np.ndarray:
    """interp1d(..., bounds_error=False, fill_value=(y[0], y[-1])) (_interpolate.py:561-574)."""
    out = call_linear(x, y, x_new)
    out[x_new < x[0]] = y[0]
    out[x_new > x[-1]] = y[-1]
    return out


def _batch_data(data: np.ndarray) -> np.ndarray:
    """A trajectory as BatchReplayEntity.add_entities holds it (entity/batch.py:88-96)."""
    d = np.nan_to_num(data)
    if d.shape[0] == 1:
        d = np.repeat(d, 2, axis=0)
        d[-1, 0] += 1e-1
    return d


def build_union_times(trajs: Sequence[np.ndarray]) -> np.ndarray:
    """The sorted union of all control-point times (entity/batch.py:98-99)."""
    cols = []
    for data in trajs:
        t = np.asarray(data, dtype=np.float64)[:, 0]
        if not np.isfinite(t).all():
            t = np.nan_to_num(t)
        if t.shape[0] == 1:  # (a single control point is held twice, 0.1 s apart: _batch_data)
            t = np.array([t[0], t[0] + 1e-1])
        cols.append(t)
    return np.unique(np.concatenate(cols)) if cols else np.zeros(0)


def build_union_table(trajs: Sequence[np.ndarray]):
    """
    Restates ``BatchReplayEntity.add_entities`` (entity/batch.py:80-112): every
    trajectory is resampled (clamped) at the sorted union of all control-point
    times.  Returns ``ts (Nk,)`` and ``X (Nk, n_ents, 6)``.
    """
    datas = [_batch_data(data) for data in trajs]
    ts = np.array(sorted(set(t for d in datas for t in d[:, 0])))
    X = np.stack([call_linear_clamped(d[:, 0], d[:, 1:], ts) for d in datas], axis=1)
    return ts, X


@dataclass
class SlotSpec:
    """One entity slot of a scenario."""

    kind: int
    traj: np.ndarray  # (K, 7) [t, x, y, z, h, p, r] as held by Trajectory.data
    box: Sequence[float] = (2.0, 4.0, 0.0, 0.0)  # width, length, center_x, center_y
    etype: int = abi.ETYPE_VEHICLE
    # VehicleController(max_steer, max_accel, max_speed, allow_reverse) of this slot, or None: SgParams
    veh_limits: Optional[Sequence[float]] = None
    speed_desired: float = 0.0
    route: Optional[np.ndarray] = None  # (R, 2)
    ref: str = ""


@dataclass
class ScenarioSpec:
    """
    One scenario: ``slots`` must be ordered like ``state.poses`` of the reference,
    i.e. agents (scenario order) first, then replayed entities
    (scenario_gym.py:232-245).
    """

    slots: List[SlotSpec]
    ego_slot: int = 0
    first_slot: int = 0
    t0: Optional[float] = None
    length: Optional[float] = None
    name: str = ""
    road_network: Optional[object] = None  # scenario_gym_b200.road_network.RoadNetwork (or None: empty)

    def finalize(self):
        if self.length is None:  # Scenario.length, scenario/scenario.py:88-91
            self.length = max(float(s.traj[:, 0].max()) for s in self.slots)
        if self.t0 is None:  # ScenarioGym.get_start_time, scenario_gym.py:213-215
            self.t0 = max(0.0, float(self.slots[self.ego_slot].traj[:, 0].min()))


@dataclass
class PackedScene:
    """numpy arrays behind ``SgScene`` (host side)."""

    N: int
    M: int
    kind: np.ndarray
    etype: np.ndarray
    box: np.ndarray
    traj_off: np.ndarray
    traj_rows: np.ndarray
    union_off: np.ndarray
    union_t: np.ndarray
    union_x: Optional[np.ndarray]  # (rows, 6, M); None: not built yet (on the device, or on first host access)
    t0: np.ndarray
    length: np.ndarray
    ego_slot: np.ndarray
    first_slot: np.ndarray
    ped_speed_desired: np.ndarray
    route_off: np.ndarray
    route_xy: np.ndarray
    n_entities: np.ndarray = field(default=None)
    # road-network surfaces (SgScene.rn_*): networks are shared between scenarios
    rn_of: np.ndarray = field(default=None)
    rn_poly_off: np.ndarray = field(default=None)
    rn_edge_off: np.ndarray = field(default=None)
    rn_edges: np.ndarray = field(default=None)
    rn_has_area: np.ndarray = field(default=None)
    veh_limits: np.ndarray = field(default=None)  # (4, N*M) per-agent VehicleController limits, or None

    def __post_init__(self):
        if self.rn_of is None:  # no road networks
            self.rn_of = np.full(self.N, -1, np.int32)
            self.rn_poly_off = np.zeros(1, np.int64)
            self.rn_edge_off = np.zeros(1, np.int64)
            self.rn_edges = np.zeros((0, 4), np.float64)
            self.rn_has_area = np.zeros(0, np.uint8)

    @property
    def n_networks(self) -> int:
        return (len(self.rn_poly_off) - 1) // 3

    @property
    def W(self) -> int:
        return (self.M + 31) // 32

    def arrays(self, union_rows: bool = True):
        """Every SgScene array that is present (optional ones -- veh_limits -- are left out when None);
        ``union_rows=False`` leaves the union table out (it is built on the device)."""
        return {k: getattr(self, k) for k in abi.SCENE_FIELDS
                if (union_rows or k != "union_x") and getattr(self, k) is not None}

    def kind_mask(self) -> int:
        """SgScene.kind_mask: OR of (1 << kind), or 0 when the vehicle fast path's promise
        (every vehicle slot present at reset) does not hold."""
        kinds = np.unique(self.kind)
        mask = 0
        for k in kinds:
            mask |= 1 << int(k)
        if mask & ~((1 << abi.KIND_EMPTY) | (1 << abi.KIND_VEHICLE)) == 0 and mask & (1 << abi.KIND_VEHICLE):
            veh = np.nonzero(self.kind == abi.KIND_VEHICLE)[0]
            first = self.traj_rows[self.traj_off[veh], 0]
            last = self.traj_rows[self.traj_off[veh + 1] - 1, 0]
            t0 = np.repeat(self.t0, self.M)[veh]
            single = (self.traj_off[veh + 1] - self.traj_off[veh]) == 1
            if not np.all(single | ((first <= t0) & (t0 <= last))):
                return 0
        return mask

    def scene_flags(self) -> int:
        """SgScene.scene_flags: facts the kernels' dispatcher needs (boxes without area)."""
        live = self.kind != abi.KIND_EMPTY
        flat = bool(np.any((self.box[0] * self.box[1] == 0.0) & live))
        return abi.SCENE_FLAT_BOXES if flat else 0

    def nbytes(self) -> int:
        return int(sum(a.nbytes for a in self.arrays().values()))

    @property
    def union_ready(self) -> bool:
        """The union table exists on the host (False: knot times only)."""
        return self._union_x is not None

    def union_on_device_ok(self) -> bool:
        """sg_build_union_x reproduces the host table: finite control points (nan_to_num is the identity)."""
        return bool(self.union_t.size) and bool(np.isfinite(self.traj_rows).all())


def _host_union_x(scene: "PackedScene") -> np.ndarray:
    """The union table of a packed scene from its control points and knot times (host build)."""
    M = scene.M
    X = np.zeros((scene.union_t.shape[0], 6, M), np.float64)
    for n in range(scene.N):
        a, b = int(scene.union_off[n]), int(scene.union_off[n + 1])
        if a == b:
            continue
        ts = scene.union_t[a:b]
        for s in np.nonzero(scene.kind[n * M:(n + 1) * M] == abi.KIND_REPLAY)[0]:
            i = n * M + int(s)
            d = _batch_data(scene.traj_rows[scene.traj_off[i]:scene.traj_off[i + 1]])
            X[a:b, :, s] = call_linear_clamped(d[:, 0], d[:, 1:], ts)
    return X


def _get_union_x(self):
    if self._union_x is None:
        self._union_x = _host_union_x(self)
    return self._union_x


def _set_union_x(self, value):
    self._union_x = value


# `union_x` may be left unbuilt (pack_scenarios(..., union_rows=False)): the engine then builds it on the
# device (sg_build_union_x); host code that reads the attribute (the oracle, tests) gets it built here
PackedScene.union_x = property(_get_union_x, _set_union_x)


def pack_road_networks(networks: Sequence[Optional[object]]):
    """
    Lower the road networks of a batch (one entry per scenario, ``None`` = empty network) to the
    polygon soups of ``SgScene.rn_*``: (rn_of, rn_poly_off, rn_edge_off, rn_edges, rn_has_area).
    Networks are stored once per distinct object; networks without any geometry count as empty.
    """
    index, uniq = {}, []
    rn_of = np.full(len(networks), -1, np.int32)
    for n, rn in enumerate(networks):
        if rn is None or not any(len(sf) for sf in rn.surfaces()):
            continue
        if id(rn) not in index:
            index[id(rn)] = len(uniq)
            uniq.append(rn)
        rn_of[n] = index[id(rn)]
    poly_off, edge_off, edges, has_area = [0], [0], [], []
    for rn in uniq:
        for surface in rn.surfaces():
            for poly in surface.polygons:
                e = poly.edges()
                edges.append(e)
                edge_off.append(edge_off[-1] + len(e))
            poly_off.append(len(edge_off) - 1)
            has_area.append(1 if surface.area > 0 else 0)
    return (rn_of, np.array(poly_off, np.int64), np.array(edge_off, np.int64),
            np.ascontiguousarray(np.concatenate(edges, axis=0) if edges else np.zeros((0, 4)), np.float64),
            np.array(has_area, np.uint8))


def pack_scenarios(specs: Sequence[ScenarioSpec], n_slots: Optional[int] = None,
                   union_rows: bool = True) -> PackedScene:
    """Pack scenario specs into one PackedScene (slots padded with EMPTY)."""
    N = len(specs)
    for s in specs:
        s.finalize()
    M = n_slots or max(len(s.slots) for s in specs)
    if M > 1024 or M < 1:
        raise ValueError("1 <= n_slots <= 1024")
    if any(len(s.slots) > M for s in specs):
        raise ValueError("scenario has more entities than n_slots")
    NM = N * M
    kind = np.zeros(NM, np.uint8)
    etype = np.zeros(NM, np.uint8)
    box = np.zeros((4, NM), np.float64)
    box[0:2] = 1.0
    speed_desired = np.zeros(NM, np.float64)
    traj_off = np.zeros(NM + 1, np.int64)
    route_off = np.zeros(NM + 1, np.int64)
    rows, routes = [], []
    union_off = np.zeros(N + 1, np.int64)
    union_t, union_x = [], []
    nrow = nroute = 0
    limits = None
    if any(sl.veh_limits is not None for sp in specs for sl in sp.slots):
        limits = np.zeros((4, NM), np.float64)
        limits[0], limits[1], limits[2] = 0.7, 5.0, np.nan  # VehicleController defaults
    for n, sp in enumerate(specs):
        replay_idx = []
        for s in range(M):
            i = n * M + s
            if s < len(sp.slots):
                sl = sp.slots[s]
                tr = np.ascontiguousarray(sl.traj, dtype=np.float64)
                if tr.ndim != 2 or tr.shape[1] != 7 or tr.shape[0] < 1:
                    raise ValueError("trajectory must be (K>=1, 7)")
                kind[i] = sl.kind
                etype[i] = sl.etype
                box[:, i] = sl.box
                speed_desired[i] = sl.speed_desired
                if limits is not None and sl.veh_limits is not None:
                    ms, ma, mv, rev = sl.veh_limits
                    limits[:, i] = (ms, ma, np.nan if mv is None else mv, 1.0 if rev else 0.0)
                rows.append(tr)
                nrow += tr.shape[0]
                if sl.route is not None:
                    r = np.ascontiguousarray(sl.route, dtype=np.float64).reshape(-1, 2)
                    routes.append(r)
                    nroute += r.shape[0]
                if sl.kind == abi.KIND_REPLAY:
                    replay_idx.append(s)
            traj_off[i + 1] = nrow
            route_off[i + 1] = nroute
        if replay_idx and not union_rows:  # knot times only: the rows are built on the device
            ts = build_union_times([sp.slots[s].traj for s in replay_idx])
            union_t.append(ts)
            union_off[n + 1] = union_off[n] + len(ts)
        elif replay_idx:
            ts, X = build_union_table([sp.slots[s].traj for s in replay_idx])
            full = np.zeros((len(ts), 6, M), np.float64)
            full[:, :, replay_idx] = np.transpose(X, (0, 2, 1))
            union_t.append(ts)
            union_x.append(full)
            union_off[n + 1] = union_off[n] + len(ts)
        else:
            union_off[n + 1] = union_off[n]
    return PackedScene(
        N=N,
        M=M,
        kind=kind,
        etype=etype,
        box=box,
        traj_off=traj_off,
        traj_rows=np.concatenate(rows, axis=0) if rows else np.zeros((0, 7)),
        union_off=union_off,
        union_t=np.concatenate(union_t) if union_t else np.zeros(0),
        union_x=None if (not union_rows and union_t) else (np.concatenate(union_x, axis=0) if union_x else np.zeros((0, 6, M))),
        t0=np.array([s.t0 for s in specs], np.float64),
        length=np.array([s.length for s in specs], np.float64),
        ego_slot=np.array([s.ego_slot for s in specs], np.int32),
        first_slot=np.array([s.first_slot for s in specs], np.int32),
        ped_speed_desired=speed_desired,
        route_off=route_off,
        route_xy=np.concatenate(routes, axis=0) if routes else np.zeros((0, 2)),
        n_entities=np.array([len(s.slots) for s in specs], np.int32),
        veh_limits=limits,
        **dict(zip(("rn_of", "rn_poly_off", "rn_edge_off", "rn_edges", "rn_has_area"),
                   pack_road_networks([s.road_network for s in specs]))),
    )


def tile_scene(scene: PackedScene, reps: int) -> PackedScene:
    """Replicate every scenario ``reps`` times round-robin (config C2: bit-identical copies)."""
    N, M = scene.N, scene.M
    order = np.tile(np.arange(N), reps)

    def plane(a):  # (..., N*M)
        lead = a.shape[:-1]
        return a.reshape(lead + (N, M))[..., order, :].reshape(lead + (len(order) * M,)).copy()

    def csr(off, data, per):
        cnt = np.diff(off).reshape(N, per)
        new_cnt = cnt[order].reshape(-1)
        new_off = np.concatenate([[0], np.cumsum(new_cnt)]).astype(np.int64)
        starts = off[:-1].reshape(N, per)[order].reshape(-1)
        idx = np.concatenate(
            [np.arange(s, s + c) for s, c in zip(starts, new_cnt)] or [np.zeros(0, np.int64)]
        ).astype(np.int64)
        return new_off, data[idx]

    traj_off, traj_rows = csr(scene.traj_off, scene.traj_rows, M)
    route_off, route_xy = csr(scene.route_off, scene.route_xy, M)
    union_off, union_t = csr(scene.union_off, scene.union_t, 1)
    union_x = csr(scene.union_off, scene.union_x, 1)[1] if scene.union_ready else None
    return PackedScene(
        N=len(order),
        M=M,
        kind=plane(scene.kind),
        etype=plane(scene.etype),
        box=plane(scene.box),
        traj_off=traj_off,
        traj_rows=traj_rows,
        union_off=union_off,
        union_t=union_t,
        union_x=union_x,
        t0=scene.t0[order].copy(),
        length=scene.length[order].copy(),
        ego_slot=scene.ego_slot[order].copy(),
        first_slot=scene.first_slot[order].copy(),
        ped_speed_desired=plane(scene.ped_speed_desired),
        route_off=route_off,
        route_xy=route_xy,
        n_entities=scene.n_entities[order].copy(),
        rn_of=scene.rn_of[order].copy(), rn_poly_off=scene.rn_poly_off, rn_edge_off=scene.rn_edge_off,
        rn_edges=scene.rn_edges, rn_has_area=scene.rn_has_area,
        veh_limits=None if scene.veh_limits is None else plane(scene.veh_limits),
    )


def slice_scene(scene: PackedScene, lo: int, hi: int) -> PackedScene:
    """Scenarios [lo, hi) of a packed scene (used to shard a batch across ranks / workers)."""
    N, M = scene.N, scene.M
    hi = min(hi, N)
    n = hi - lo

    def plane(a):
        lead = a.shape[:-1]
        return np.ascontiguousarray(a.reshape(lead + (N, M))[..., lo:hi, :].reshape(lead + (n * M,)))

    def csr(off, data, per):
        a, b = off[lo * per], off[hi * per]
        return (off[lo * per: hi * per + 1] - a).astype(np.int64), np.ascontiguousarray(data[a:b])

    traj_off, traj_rows = csr(scene.traj_off, scene.traj_rows, M)
    route_off, route_xy = csr(scene.route_off, scene.route_xy, M)
    union_off, union_t = csr(scene.union_off, scene.union_t, 1)
    union_x = csr(scene.union_off, scene.union_x, 1)[1] if scene.union_ready else None
    return PackedScene(
        N=n, M=M, kind=plane(scene.kind), etype=plane(scene.etype), box=plane(scene.box),
        traj_off=traj_off, traj_rows=traj_rows, union_off=union_off, union_t=union_t,
        union_x=union_x, t0=scene.t0[lo:hi].copy(), length=scene.length[lo:hi].copy(),
        ego_slot=scene.ego_slot[lo:hi].copy(), first_slot=scene.first_slot[lo:hi].copy(),
        ped_speed_desired=plane(scene.ped_speed_desired), route_off=route_off, route_xy=route_xy,
        n_entities=scene.n_entities[lo:hi].copy(),
        rn_of=scene.rn_of[lo:hi].copy(), rn_poly_off=scene.rn_poly_off, rn_edge_off=scene.rn_edge_off,
        rn_edges=scene.rn_edges, rn_has_area=scene.rn_has_area,
        veh_limits=None if scene.veh_limits is None else plane(scene.veh_limits),
    )
