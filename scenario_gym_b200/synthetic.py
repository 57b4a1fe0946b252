"""
Synthetic scenario generators for the configurations named in BASELINE.json
(SURVEY.md section 8d): C3 random-action vehicles, C4 social-force crowds,
C5 dense highway with RSS.  Everything is drawn from ``numpy.random.default_rng``
so the reference (golden generation), the CPU oracle and the GPU engine consume
bit-identical inputs.  Scenes are packed vectorised (no per-entity Python objects)
so N = 100k x M = 64 packs in seconds.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import abi
from .action_rng import ActionRng
from .packing import PackedScene

# tests/input_files/Catalogs/Scenario_Gym/VehicleCatalogs/ScenarioGymVehicleCatalog.xosc:8-9 (car1)
CAR1_BOX = (2.0, 4.2, 1.37, 0.0)  # width, length, center_x, center_y
# .../PedestrianCatalogs/ScenarioGymPedestrianCatalog.xosc:7-8 (pedestrian1)
PED1_BOX = (0.69, 0.7, 0.0, 0.0)


@dataclass
class SyntheticConfig:
    """Plain arrays describing N scenarios x M entities."""

    name: str
    N: int
    M: int
    T: int
    dt: float
    x0: np.ndarray  # (N, M)
    y0: np.ndarray
    h0: np.ndarray
    v0: np.ndarray
    box: np.ndarray  # (4,) or (N, M, 4)
    kind: np.ndarray  # (N, M) SgKind
    etype: np.ndarray  # (N, M)
    actions: Optional[np.ndarray] = None  # (T, 2, N*M) accel, steer
    speed_desired: Optional[np.ndarray] = None  # (N, M)
    goal: Optional[np.ndarray] = None  # (N, M, 2) pedestrians' route end point
    # the action table as a description of its place in the numpy stream (device-side action source);
    # always set for the random-action configs, `actions` only when the table was materialised
    action_rng: Optional[ActionRng] = None

    @property
    def t_end(self) -> float:
        # scenario.length chosen so that exactly T ticks run: done when t + dt > length
        return (self.T + 0.5) * self.dt


def two_knot_rows(cfg: SyntheticConfig) -> np.ndarray:
    """Each entity's trajectory: straight line at speed v0 along h0, knots at 0 and t_end."""
    N, M = cfg.N, cfg.M
    rows = np.zeros((N * M, 2, 7))
    te = cfg.t_end
    x0, y0, h0, v0 = (a.reshape(-1) for a in (cfg.x0, cfg.y0, cfg.h0, cfg.v0))
    rows[:, 0, 1], rows[:, 0, 2], rows[:, 0, 4] = x0, y0, h0
    rows[:, 1, 0] = te
    rows[:, 1, 1] = x0 + v0 * np.cos(h0) * te
    rows[:, 1, 2] = y0 + v0 * np.sin(h0) * te
    rows[:, 1, 4] = h0
    return rows


def pack_synthetic(cfg: SyntheticConfig, road_network=None) -> PackedScene:
    """Pack a synthetic configuration; `road_network` (optional) is shared by all its scenarios."""
    from .packing import pack_road_networks

    N, M = cfg.N, cfg.M
    NM = N * M
    rows = two_knot_rows(cfg).reshape(NM * 2, 7)
    box = np.empty((4, NM))
    if cfg.box.ndim == 1:
        box[:] = cfg.box[:, None]
    else:
        box[:] = cfg.box.reshape(NM, 4).T
    if cfg.goal is not None:
        route = np.stack(
            [np.stack([cfg.x0, cfg.y0], -1).reshape(NM, 2), cfg.goal.reshape(NM, 2)], axis=1
        ).reshape(NM * 2, 2)
        route_off = np.arange(NM + 1, dtype=np.int64) * 2
        is_ped = cfg.kind.reshape(-1) == abi.KIND_PEDESTRIAN
        if not is_ped.all():  # non-pedestrians have no route
            cnt = np.where(is_ped, 2, 0)
            route_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
            route = route.reshape(NM, 2, 2)[is_ped].reshape(-1, 2)
    else:
        route = np.zeros((0, 2))
        route_off = np.zeros(NM + 1, np.int64)
    if (cfg.kind == abi.KIND_REPLAY).any():
        raise ValueError("synthetic fast path has no batch-replay entities")
    return PackedScene(
        N=N,
        M=M,
        kind=np.ascontiguousarray(cfg.kind.reshape(-1), np.uint8),
        etype=np.ascontiguousarray(cfg.etype.reshape(-1), np.uint8),
        box=box,
        traj_off=np.arange(NM + 1, dtype=np.int64) * 2,
        traj_rows=rows,
        union_off=np.zeros(N + 1, np.int64),
        union_t=np.zeros(0),
        union_x=np.zeros((0, 6, M)),
        t0=np.zeros(N),
        length=np.full(N, cfg.t_end),
        ego_slot=np.zeros(N, np.int32),
        first_slot=np.zeros(N, np.int32),
        ped_speed_desired=(
            np.ascontiguousarray(cfg.speed_desired.reshape(-1))
            if cfg.speed_desired is not None
            else np.zeros(NM)
        ),
        route_off=route_off,
        route_xy=np.ascontiguousarray(route),
        n_entities=np.full(N, M, np.int32),
        **(dict(zip(("rn_of", "rn_poly_off", "rn_edge_off", "rn_edges", "rn_has_area"),
                    pack_road_networks([road_network] * N))) if road_network is not None else {}),
    )


def _uniform_into(rng, out: np.ndarray, low: float, high: float) -> None:
    """Same stream and bits as rng.uniform(low, high, out.shape), written into `out`."""
    rng.random(out=out)
    out *= high - low
    out += low


def vehicles_config(seed: int, N: int, M: int = 64, T: int = 256, dt: float = 0.1,
                    half_extent: float = 200.0, name: str = "C3",
                    actions_out: Optional[np.ndarray] = None, materialise: bool = True) -> SyntheticConfig:
    """
    C3: M VehicleController agents per scenario with random accel/steer actions.
    x,y ~ U(-half_extent, half_extent), h ~ U(-pi, pi), v0 ~ U(0, 15),
    accel ~ U(-6, 6) (exercises the +-5 clip), steer ~ U(-1, 1) (+-0.7 clip).
    The (T, 2, N*M) action table follows the initial conditions in the generator's stream, plane by
    plane (all accel rows, then all steer rows); ``materialise=False`` only describes it
    (``action_rng``) for the device-side action source.
    """
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-half_extent, half_extent, (N, M))
    y0 = rng.uniform(-half_extent, half_extent, (N, M))
    h0 = rng.uniform(-np.pi, np.pi, (N, M))
    v0 = rng.uniform(0.0, 15.0, (N, M))
    planes = ((-6.0, 6.0), (-1.0, 1.0))
    arng = ActionRng.from_generator(rng, offset=(0, T * N * M), tick_stride=N * M,
                                    low=[lo for lo, _ in planes], high=[hi for _, hi in planes],
                                    n_ticks=T, nm=N * M)
    actions = None
    if materialise or actions_out is not None:
        actions = np.empty((T, 2, N * M)) if actions_out is None else actions_out
        assert actions.shape == (T, 2, N * M) and actions.dtype == np.float64
        # drawn in place: the table can be a view of pinned host memory
        tmp = np.empty(N * M)
        for c, (lo, hi) in enumerate(planes):
            for k in range(T):
                _uniform_into(rng, tmp, lo, hi)
                actions[k, c] = tmp
    return SyntheticConfig(
        name=name, N=N, M=M, T=T, dt=dt, x0=x0, y0=y0, h0=h0, v0=v0,
        box=np.array(CAR1_BOX), kind=np.full((N, M), abi.KIND_VEHICLE, np.uint8),
        etype=np.full((N, M), abi.ETYPE_VEHICLE, np.uint8), actions=actions, action_rng=arng,
    )


def highway_config(seed: int, N: int, M: int = 256, T: int = 256, dt: float = 0.1,
                   lanes: int = 4, name: str = "C5", materialise: bool = True) -> SyntheticConfig:
    """
    C5: `lanes` lanes x (M / lanes) vehicles, lane width 3.7 m, headway U(8, 40) m,
    heading ~ 0 +- 0.02, v0 ~ U(20, 35); small random actions.  Ego = slot 0.
    """
    rng = np.random.default_rng(seed)
    per = M // lanes
    assert per * lanes == M
    headway = rng.uniform(8.0, 40.0, (N, lanes, per))
    x0 = np.cumsum(headway, axis=2).reshape(N, M)
    y0 = np.repeat(np.arange(lanes) * 3.7, per)[None, :] + rng.uniform(-0.3, 0.3, (N, M))
    h0 = rng.uniform(-0.02, 0.02, (N, M))
    v0 = rng.uniform(20.0, 35.0, (N, M))
    # put the ego (slot 0) in the middle of the pack so it has traffic on all sides
    mid = per // 2
    for a in (x0, y0, h0, v0):
        a[:, [0, mid]] = a[:, [mid, 0]]
    arng = ActionRng.from_generator(rng, offset=(0, T * N * M), tick_stride=N * M, low=(-2.0, -0.02),
                                    high=(2.0, 0.02), n_ticks=T, nm=N * M)
    actions = None
    if materialise:
        actions = np.empty((T, 2, N * M))
        actions[:, 0] = rng.uniform(-2.0, 2.0, (T, N * M))
        actions[:, 1] = rng.uniform(-0.02, 0.02, (T, N * M))
    return SyntheticConfig(
        name=name, N=N, M=M, T=T, dt=dt, x0=x0, y0=y0, h0=h0, v0=v0,
        box=np.array(CAR1_BOX), kind=np.full((N, M), abi.KIND_VEHICLE, np.uint8),
        etype=np.full((N, M), abi.ETYPE_VEHICLE, np.uint8), actions=actions, action_rng=arng,
    )


def crowd_config(seed: int, N: int, M: int = 1024, T: int = 128, dt: float = 1.0 / 15.0,
                 side: float = 40.0, name: str = "C4") -> SyntheticConfig:
    """
    C4: slot 0 is an ego vehicle replaying a straight trajectory outside the square;
    slots 1.. are social-force pedestrians starting uniformly in a side x side square
    with a 2-point route to a uniformly drawn goal; speed_desired ~ U(0.5, 1.5) * 1.4.
    """
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(0.0, side, (N, M))
    y0 = rng.uniform(0.0, side, (N, M))
    goal = rng.uniform(0.0, side, (N, M, 2))
    d = goal - np.stack([x0, y0], -1)
    # initial heading / speed deliberately differ from the goal direction / desired speed:
    # otherwise the goal force cancels to rounding noise and the heading atan2(F) of a
    # neighbour-free pedestrian is ill-conditioned (no parity statement is possible)
    h0 = np.arctan2(d[..., 1], d[..., 0]) + rng.uniform(-0.5, 0.5, (N, M))
    speed_desired = rng.uniform(0.5, 1.5, (N, M)) * 1.4
    v0 = speed_desired * rng.uniform(0.3, 0.9, (N, M))
    kind = np.full((N, M), abi.KIND_PEDESTRIAN, np.uint8)
    etype = np.full((N, M), abi.ETYPE_PEDESTRIAN, np.uint8)
    box = np.empty((N, M, 4))
    box[:] = PED1_BOX
    # ego vehicle driving along y = -10
    kind[:, 0] = abi.KIND_AGENT_REPLAY
    etype[:, 0] = abi.ETYPE_VEHICLE
    box[:, 0] = CAR1_BOX
    x0[:, 0], y0[:, 0], h0[:, 0], v0[:, 0] = 0.0, -10.0, 0.0, 5.0
    return SyntheticConfig(
        name=name, N=N, M=M, T=T, dt=dt, x0=x0, y0=y0, h0=h0, v0=v0, box=box, kind=kind,
        etype=etype, speed_desired=speed_desired, goal=goal,
    )
