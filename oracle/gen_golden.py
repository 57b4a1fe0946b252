"""
TEST INFRASTRUCTURE ONLY -- generate golden vectors by executing the UNMODIFIED
reference (driskai/scenario_gym at /root/reference) in the authoring container.

    python -m oracle.gen_golden            # writes tests/golden/*.npz + manifest.json

The reference is imported through ``oracle/refshim`` (stubs for its absent
third-party roots; Shapely calls go to the restated subset).  Outputs are what the
reference computes through its own public API: ``ScenarioGym.rollout`` /
``step``, ``state.poses/velocities/distances/collisions()``, metric ``get_state()``,
``RSSDistances`` attributes, ``PedestrianAgent.force``.

Cases
  xosc   the 23 OpenSCENARIO files of tests/input_files/Scenarios, default gym
         (timestep 1/30, max_length, default create_agent, relabel True) [C1/C2]
  xosc_norelabel / xosc_persist   two files with relabel=False / persist=True
  veh    VehicleController agents with pre-drawn actions, dense so they collide [C3]
  rss    highway-like traffic with RSSDistances + RSS [C5]
  ped    social-force pedestrians (std_lon = std_lat = 0) + replayed ego [C4]
  unit   Trajectory / bounding-box known answers
  future FutureCollisionDetector flags on the test scenarios (every 20th tick, two horizons)
"""
from __future__ import annotations

import glob
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import refshim  # noqa: E402

refshim.install()

import scenario_gym as ref  # noqa: E402
from scenario_gym.action import VehicleAction  # noqa: E402
from scenario_gym.agent import Agent  # noqa: E402
from scenario_gym.catalog_entry import BoundingBox, CatalogEntry  # noqa: E402
from scenario_gym.controller import VehicleController  # noqa: E402
from scenario_gym.entity import Entity, Pedestrian, Vehicle  # noqa: E402
from scenario_gym.metrics import (  # noqa: E402
    RSS,
    CollisionMetric,
    EgoAvgSpeed,
    EgoDistanceTravelled,
    EgoMaxSpeed,
    Metric,
    RSSDistances,
)
from scenario_gym.pedestrian.agent import PedestrianAgent  # noqa: E402
from scenario_gym.pedestrian.social_force import SocialForce, SocialForceParameters  # noqa: E402
from scenario_gym.scenario import Scenario  # noqa: E402
from scenario_gym.scenario_gym import ScenarioGym  # noqa: E402
from scenario_gym.sensor import EgoLocalizationSensor  # noqa: E402
from scenario_gym.trajectory import Trajectory  # noqa: E402
from scenario_gym.xosc_interface import import_scenario  # noqa: E402

from scenario_gym_b200 import abi, synthetic  # noqa: E402
from oracle import golden_cases  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")
SCEN_DIR = os.path.join(refshim.REFERENCE_ROOT, "tests", "input_files", "Scenarios")

RSS_CODE = {v: k for k, v in abi.RSS_RECORD_NAMES.items()}


def etype_of(e) -> int:
    if isinstance(e, Vehicle):
        return abi.ETYPE_VEHICLE
    if isinstance(e, Pedestrian) or e.type == "Pedestrian":
        return abi.ETYPE_PEDESTRIAN
    return abi.ETYPE_MISC


class Recorder(Metric):
    """Reference-side metric that snapshots the public state every tick."""

    name = "recorder"

    def __init__(self, rss_cb=None, ped_agents=None):
        super().__init__()
        self.rss_cb = rss_cb
        self.ped_agents = ped_agents

    def _reset(self, state):
        self.ents = state.scenario.entities
        self.idx = {e: i for i, e in enumerate(self.ents)}
        self.ego = state.scenario.ego
        self.t, self.present, self.pose, self.vel, self.dist = [], [], [], [], []
        self.pairs, self.ego_events, self.last = [], [], []
        self.rss_sd, self.rss_ratio, self.rss_rec, self._rss_len = [], [], [], None
        self.force, self.goal = [], []
        self.tick = 0
        self._snap(state)

    def _snap(self, state):
        M = len(self.ents)
        pres = np.zeros(M, np.uint8)
        pose = np.full((M, 6), np.nan)
        vel = np.full((M, 6), np.nan)
        dist = np.zeros(M)
        for e, i in self.idx.items():
            dist[i] = state.distances[e]
            if e in state.poses:
                pres[i] = 1
                pose[i] = state.poses[e]
                vel[i] = state.velocities[e]
        self.t.append(state.t)
        self.present.append(pres)
        self.pose.append(pose)
        self.vel.append(vel)
        self.dist.append(dist)
        if self.rss_cb is not None:
            cb = self.rss_cb
            sd = np.zeros((M, 2))
            ratio = np.full((M, 2), np.inf)
            rec = np.full(M, abi.RSS_NONE, np.uint8)
            if self._rss_len is None:
                self._rss_len = {e: 1 for e in cb.intersect}
            for e, i in self.idx.items():
                if e in cb.safe_distances:
                    sd[i] = cb.safe_distances[e]
                ratio[i] = cb.entity_safe_ratios[e]
                if e in cb.intersect and len(cb.intersect[e]) > self._rss_len[e]:
                    last = cb.intersect[e][-1]
                    rec[i] = RSS_CODE["found"] if isinstance(last, list) else RSS_CODE[last]
                    self._rss_len[e] = len(cb.intersect[e])
            self.rss_sd.append(sd)
            self.rss_ratio.append(ratio)
            self.rss_rec.append(rec)
        if self.ped_agents is not None:
            f = np.zeros((M, 2))
            g = np.zeros(M, np.int32)
            for e, a in self.ped_agents.items():
                f[self.idx[e]] = a.force
                g[self.idx[e]] = a.goal_idx
            self.force.append(f)
            self.goal.append(g)

    def _step(self, state):
        self.tick += 1
        self._snap(state)
        coll = state.collisions()
        seen = set()
        for e, others in coll.items():
            for o in others:
                a, b = sorted((self.idx[e], self.idx[o]))
                seen.add((a, b))
        for a, b in sorted(seen):
            self.pairs.append((self.tick, a, b))
        if self.ego in coll:
            now = coll[self.ego]
            for o in now:
                if o not in self.last:
                    self.ego_events.append((self.tick, self.idx[o], state.t))
            self.last = list(now)

    def get_state(self):
        return None

    def dump(self, dec: int):
        T = len(self.t)
        keep = sorted(set(range(0, T, dec)) | {T - 1, 1 if T > 1 else 0})
        pose = np.array(self.pose)
        out = dict(
            t=np.array(self.t),
            present=np.array(self.present),
            keep=np.array(keep, np.int32),
            pose=pose[keep],
            vel=np.array(self.vel)[keep],
            dist=np.array(self.dist)[keep],
            pose_sum=np.nansum(pose, axis=1),
            pairs=np.array(self.pairs, np.int32).reshape(-1, 3),
            ego_events=np.array([(a, b) for a, b, _ in self.ego_events], np.int32).reshape(-1, 2),
            ego_event_t=np.array([c for _, _, c in self.ego_events], np.float64),
        )
        if self.rss_cb is not None:
            out["rss_sd"] = np.array(self.rss_sd)
            out["rss_ratio"] = np.array(self.rss_ratio)
            out["rss_rec"] = np.array(self.rss_rec)
        if self.ped_agents is not None:
            out["force"] = np.array(self.force)
            out["goal"] = np.array(self.goal)
        return out


def scenario_inputs(scenario, agents_idx):
    """Inputs of a scenario as plain arrays (what the engine packs)."""
    ents = scenario.entities
    out = {
        "n_entities": np.int32(len(ents)),
        "box": np.array(
            [[e.bounding_box.width, e.bounding_box.length, e.bounding_box.center_x,
              e.bounding_box.center_y] for e in ents]
        ),
        "etype": np.array([etype_of(e) for e in ents], np.uint8),
        "is_agent": np.array([i in agents_idx for i in range(len(ents))], np.uint8),
        "ego": np.int32(ents.index(scenario.ego)),
    }
    for i, e in enumerate(ents):
        out[f"traj{i}"] = np.array(e.trajectory.data)
    return out


def run_gym(gym, rec, dec):
    gym.rollout()
    m = gym.get_metrics()
    out = rec.dump(dec)
    out["n_ticks"] = np.int32(rec.tick)
    for k in ("ego_avg_speed", "ego_max_speed", "ego_distance_travelled"):
        if k in m:
            out[k] = np.float64(m[k])
    return out, m


def flat(prefix, d, store):
    for k, v in d.items():
        store[f"{prefix}/{k}"] = v


# --------------------------------------------------------------------------- xosc
def gen_xosc(store, manifest):
    files = sorted(glob.glob(os.path.join(SCEN_DIR, "*.xosc")))
    names = []
    for f in files:
        name = os.path.splitext(os.path.basename(f))[0]
        names.append(name)
        for variant, kw in (("", {}), ("_norelabel", {}), ("_persist", {"persist": True})):
            if variant and name[:8] not in ("a5e43fe4", "41dac6fa", "5c5188e0"):
                continue
            rec = Recorder()
            gym = ScenarioGym(
                metrics=[EgoAvgSpeed(), EgoMaxSpeed(), EgoDistanceTravelled(), rec], **kw
            )
            gym.load_scenario(f, relabel=(variant != "_norelabel"))
            sc = gym.state.scenario
            agents_idx = {sc.entities.index(e) for e in gym.state.agents}
            key = f"xosc{variant}/{name}"
            flat(key + "/in", scenario_inputs(sc, agents_idx), store)
            out, m = run_gym(gym, rec, dec=16 if not variant else 32)
            out["t_end"] = np.float64(gym.state.t)
            flat(key + "/out", out, store)
            manifest.setdefault("xosc" + variant, {})[name] = {
                "refs": [e.ref for e in sc.entities],
                "persist": bool(kw.get("persist", False)),
                "n_ticks": int(rec.tick),
            }
            print(key, len(sc.entities), rec.tick, repr(float(gym.state.t)), len(rec.pairs),
                  len(rec.ego_events))
    # the reference's CollisionMetric on a pedestrian-hazard scenario (non_vehicle path)
    f = [x for x in files if "379d4431" in x][0]
    gym = ScenarioGym(metrics=[CollisionMetric()])
    gym.load_scenario(f, relabel=True)
    gym.rollout()
    manifest["collision_metric_379d4431"] = [
        [float(t), r, c] for t, r, c in gym.get_metrics()["collisions"]
    ]


# --------------------------------------------------------------------------- synthetic
class TableVehicleAgent(Agent):
    """Reference Agent that replays a pre-drawn (accel, steer) table through VehicleController."""

    def __init__(self, entity, table):
        super().__init__(entity, VehicleController(entity), EgoLocalizationSensor(entity))
        self.table = table
        self.k = 0

    def _reset(self):
        self.k = 0

    def _step(self, observation):
        a = self.table[self.k]
        self.k += 1
        return VehicleAction(a[0], a[1])


def catalog_entry(etype, box):
    bb = BoundingBox(*[float(b) for b in box])
    ctype = {abi.ETYPE_VEHICLE: "Vehicle", abi.ETYPE_PEDESTRIAN: "Pedestrian"}.get(etype, "MiscObject")
    return CatalogEntry(None, "synthetic", "car", ctype, bb, {}, [])


def ref_scenario(cfg: synthetic.SyntheticConfig, n: int, road_network=None):
    rows = synthetic.two_knot_rows(cfg).reshape(cfg.N, cfg.M, 2, 7)
    ents = []
    for m in range(cfg.M):
        box = cfg.box if cfg.box.ndim == 1 else cfg.box[n, m]
        et = int(cfg.etype[n, m])
        ce = catalog_entry(et, box)
        Cls = {abi.ETYPE_VEHICLE: Vehicle, abi.ETYPE_PEDESTRIAN: Pedestrian}.get(et, Entity)
        e = Cls(ce, ref="ego" if m == 0 else f"entity_{m}")
        e.trajectory = Trajectory(rows[n, m])
        assert np.array_equal(e.trajectory.data, rows[n, m]), "trajectory not canonical"
        ents.append(e)
    return Scenario(ents, name=f"{cfg.name}_{n}", road_network=road_network)


def gen_vehicle_like(tag, cfg, store, manifest, rss=False, terminal=None):
    for n in range(cfg.N):
        sc = ref_scenario(cfg, n)
        acts = cfg.actions.reshape(cfg.T, 2, cfg.N, cfg.M)
        cb = RSSDistances() if rss else None
        rec = Recorder(rss_cb=cb)
        metrics = [EgoAvgSpeed(), EgoMaxSpeed(), EgoDistanceTravelled()]
        if rss:
            metrics.append(RSS())
        metrics.append(rec)
        gym = ScenarioGym(
            timestep=cfg.dt,
            metrics=metrics,
            state_callbacks=[cb] if rss else None,
            terminal_conditions=terminal,
        )

        def create_agent(scenario, entity, n=n, acts=acts):
            m = scenario.entities.index(entity)
            if cfg.kind[n, m] == abi.KIND_VEHICLE:
                return TableVehicleAgent(entity, acts[:, :, n, m])
            return None

        gym.set_scenario(sc, create_agent=create_agent)
        out, m = run_gym(gym, rec, dec=8)
        out["t_end"] = np.float64(gym.state.t)
        if rss:
            out["rss_safe_longitudinal"] = np.uint8(m["RSS_safe_longitudinal"])
            out["rss_safe_lateral"] = np.uint8(m["RSS_safe_lateral"])
        flat(f"{tag}/{n}/out", out, store)
        print(tag, n, rec.tick, repr(float(gym.state.t)), "pairs", len(rec.pairs), "events",
              len(rec.ego_events), {k: v for k, v in m.items() if k.startswith("RSS")})
    manifest[tag] = {"N": cfg.N, "M": cfg.M, "T": cfg.T, "dt": cfg.dt}


def gen_ped(store, manifest):
    cfg = golden_cases.ped_cfg()
    params = SocialForceParameters(std_lon=0.0, std_lat=0.0)
    for n in range(cfg.N):
        sc = ref_scenario(cfg, n, road_network=refshim.EmptyRoadNetwork())
        peds = {}

        def create_agent(scenario, entity, n=n, peds=peds):
            m = scenario.entities.index(entity)
            if cfg.kind[n, m] == abi.KIND_PEDESTRIAN:
                route = [np.array([cfg.x0[n, m], cfg.y0[n, m]]), np.array(cfg.goal[n, m])]
                a = PedestrianAgent(entity, route, float(cfg.speed_desired[n, m]), SocialForce(params))
                peds[entity] = a
                return a
            return ref.agent._create_agent(scenario, entity)

        rec = Recorder(ped_agents=peds)
        gym = ScenarioGym(timestep=cfg.dt, metrics=[EgoAvgSpeed(), rec])
        gym.set_scenario(sc, create_agent=create_agent)
        out, m = run_gym(gym, rec, dec=4)
        out["t_end"] = np.float64(gym.state.t)
        flat(f"ped/{n}/out", out, store)
        print("ped", n, rec.tick, repr(float(gym.state.t)), "pairs", len(rec.pairs),
              "goal reached", int((np.array(rec.goal[-1]) > 1).sum()))
    manifest["ped"] = {"N": cfg.N, "M": cfg.M, "T": cfg.T, "dt": cfg.dt}


# --------------------------------------------------------------------------- road networks
def ref_road_network(geometry):
    """The reference's own RoadNetwork over the geometry of oracle/golden_cases.py (shim polygons)."""
    from scenario_gym.road_network import Building, Pavement, Road, RoadNetwork
    from shapely.geometry import LineString, Polygon

    def poly(b):
        if isinstance(b, dict):
            return Polygon(b["exterior"], holes=b["interiors"])
        return Polygon(b)

    center = LineString([(0.0, 0.0), (1.0, 0.0)])
    return RoadNetwork(
        roads=[Road(f"road_{k}", poly(b), center, []) for k, b in enumerate(geometry["roads"])],
        intersections=[],
        pavements=[Pavement(f"pavement_{k}", poly(b), center) for k, b in enumerate(geometry["pavements"])],
        buildings=[Building(f"building_{k}", poly(b)) for k, b in enumerate(geometry["buildings"])],
    )


def gen_road(store, manifest):
    """Social-force boundary forces (buildings in the crowd) and the ego_off_road terminal condition."""
    # (a) pedestrians among buildings
    cfg = golden_cases.ped_cfg()
    rn = ref_road_network(golden_cases.ROAD_PED_GEOMETRY)
    params = SocialForceParameters(std_lon=0.0, std_lat=0.0)
    for n in range(cfg.N):
        sc = ref_scenario(cfg, n, road_network=rn)
        peds = {}

        def create_agent(scenario, entity, n=n, peds=peds):
            m = scenario.entities.index(entity)
            if cfg.kind[n, m] == abi.KIND_PEDESTRIAN:
                route = [np.array([cfg.x0[n, m], cfg.y0[n, m]]), np.array(cfg.goal[n, m])]
                a = PedestrianAgent(entity, route, float(cfg.speed_desired[n, m]), SocialForce(params))
                peds[entity] = a
                return a
            return ref.agent._create_agent(scenario, entity)

        rec = Recorder(ped_agents=peds)
        gym = ScenarioGym(timestep=cfg.dt, metrics=[EgoAvgSpeed(), rec])
        gym.set_scenario(sc, create_agent=create_agent)
        out, m = run_gym(gym, rec, dec=4)
        out["t_end"] = np.float64(gym.state.t)
        flat(f"road_ped/{n}/out", out, store)
        print("road_ped", n, rec.tick, repr(float(gym.state.t)), "pairs", len(rec.pairs))
    # (b) vehicles leaving the driveable surface
    cfg = golden_cases.veh_cfg()
    rn = ref_road_network(golden_cases.ROAD_VEH_GEOMETRY)
    ticks = []
    for n in range(cfg.N):
        sc = ref_scenario(cfg, n, road_network=rn)
        acts = cfg.actions.reshape(cfg.T, 2, cfg.N, cfg.M)
        rec = Recorder()
        gym = ScenarioGym(timestep=cfg.dt, metrics=[EgoAvgSpeed(), EgoMaxSpeed(), EgoDistanceTravelled(), rec],
                          terminal_conditions=["max_length", "ego_off_road"])

        def create_agent(scenario, entity, n=n, acts=acts):
            m = scenario.entities.index(entity)
            return TableVehicleAgent(entity, acts[:, :, n, m])

        gym.set_scenario(sc, create_agent=create_agent)
        out, m = run_gym(gym, rec, dec=8)
        out["t_end"] = np.float64(gym.state.t)
        flat(f"road_veh/{n}/out", out, store)
        ticks.append(rec.tick)
        print("road_veh", n, rec.tick, repr(float(gym.state.t)))
    # (c) the reference's own test scenarios with their JSON road networks and ego_off_road
    import xml.etree.ElementTree as ET

    from scenario_gym.road_network import RoadNetwork as RefRoadNetwork

    xticks = {}
    for path in sorted(glob.glob(os.path.join(SCEN_DIR, "*.xosc"))):
        name = os.path.splitext(os.path.basename(path))[0]
        xroot = ET.parse(path).getroot()
        node = xroot.find("RoadNetwork/SceneGraphFile")
        if node is None:
            node = xroot.find("RoadNetwork/LogicFile")
        if node is None:
            continue
        rn_file = os.path.join(os.path.dirname(path), node.attrib["filepath"])
        if os.path.splitext(rn_file)[1] == "":
            rn_file += ".json"
        if not os.path.exists(rn_file):
            continue
        sc = import_scenario(path)
        sc.road_network = RefRoadNetwork.create_from_json(rn_file)
        stem = os.path.splitext(os.path.basename(rn_file))[0].replace(" ", "_")
        store[f"road_xosc/{name}/network"] = np.array(stem)
        if f"road_net/{stem}/n" not in store:  # the network's surfaces as the reference builds them
            rn_ = sc.road_network
            counts = []
            for tag_, surf in (("d", rn_.driveable_surface), ("w", rn_.walkable_surface), ("i", rn_.impenetrable_surface)):
                counts.append(len(surf.geoms))
                for k_, g_ in enumerate(surf.geoms):
                    rings_ = g_.rings()
                    store[f"road_net/{stem}/{tag_}{k_}/ext"] = np.asarray(rings_[0], np.float64)
                    for j_, h_ in enumerate(rings_[1:]):
                        store[f"road_net/{stem}/{tag_}{k_}/hole{j_}"] = np.asarray(h_, np.float64)
            store[f"road_net/{stem}/n"] = np.array(counts, np.int64)
        gym = ScenarioGym(metrics=[EgoAvgSpeed(), EgoDistanceTravelled()],
                          terminal_conditions=["max_length", "ego_off_road"])
        gym.set_scenario(sc)
        ticks_n = 0
        gym.reset_scenario()
        while not gym.state.is_done:
            gym.step()
            ticks_n += 1
        m = gym.get_metrics()
        store[f"road_xosc/{name}/n_ticks"] = np.int64(ticks_n)
        store[f"road_xosc/{name}/t_end"] = np.float64(gym.state.t)
        store[f"road_xosc/{name}/ego_avg_speed"] = np.float64(m["ego_avg_speed"])
        store[f"road_xosc/{name}/ego_distance_travelled"] = np.float64(m["ego_distance_travelled"])
        xticks[name] = ticks_n
        print("road_xosc", name[:8], ticks_n, repr(float(gym.state.t)))
    manifest["road"] = {"ped": {"N": golden_cases.ped_cfg().N}, "veh_ticks": ticks, "xosc_ticks": xticks}


# --------------------------------------------------------------------------- PID
def gen_pid(store, manifest):
    """PIDAgent / PIDController (agent.py:131-148, controller.py:143-258) on two test scenarios."""
    from scenario_gym.agent import PIDAgent

    cases = {
        # the reference's own test, tests/test_controller.py:7-25
        "a98d5c7d": dict(timestep=0.1, kwargs=dict(accel_Kp=2.0, max_accel=5.0, max_steer=float(np.pi / 90))),
        "3fee6507": dict(timestep=1.0 / 30.0, kwargs={}),
    }
    files = sorted(glob.glob(os.path.join(SCEN_DIR, "*.xosc")))
    for short, cfg in cases.items():
        f = [x for x in files if short in x][0]
        name = os.path.splitext(os.path.basename(f))[0]
        rec = Recorder()
        gym = ScenarioGym(timestep=cfg["timestep"],
                          metrics=[EgoAvgSpeed(), EgoMaxSpeed(), EgoDistanceTravelled(), rec])

        def create_agent(s, e, kw=cfg["kwargs"]):
            if e.ref == "ego":
                return PIDAgent(e, **kw)

        gym.load_scenario(f, create_agent=create_agent, relabel=True)
        out, m = run_gym(gym, rec, dec=8)
        out["t_end"] = np.float64(gym.state.t)
        flat(f"pid/{name}/out", out, store)
        manifest.setdefault("pid", {})[name] = {"timestep": cfg["timestep"], "kwargs": cfg["kwargs"],
                                                "n_ticks": int(rec.tick)}
        print("pid", name, rec.tick, repr(float(gym.state.t)), m)


# --------------------------------------------------------------------------- unit vectors
def gen_unit(store, manifest):
    rng = np.random.default_rng(123)
    # bounding box corners (entity/base.py:100-138)
    ce = catalog_entry(abi.ETYPE_VEHICLE, synthetic.CAR1_BOX)
    e = Entity(ce, ref="x")
    poses = np.zeros((256, 6))
    poses[:, 0:2] = rng.uniform(-300, 300, (256, 2))
    poses[:, 3] = rng.uniform(-7, 7, 256)
    store["unit/box_pose"] = poses
    store["unit/box_points"] = np.array([e.get_bounding_box_points(p) for p in poses])
    # trajectory interpolation / extrapolation (trajectory.py:142-205, 243-273)
    K = 9
    data = np.zeros((K, 7))
    data[:, 0] = np.sort(rng.uniform(1.0, 9.0, K))
    data[:, 1:] = rng.normal(size=(K, 6)) * 10
    data[:, 4] = np.cumsum(rng.uniform(-0.3, 0.3, K))
    tr = Trajectory(data)
    ts = np.concatenate([rng.uniform(-2, 12, 200), data[:, 0], [data[0, 0] - 1e-9, data[-1, 0] + 1e-9]])
    store["unit/traj_data"] = np.array(tr.data)
    store["unit/traj_ts"] = ts
    for mode, ext in ((0, False), (1, (False, False)), (2, True)):
        res = [tr.position_at_t(float(t), extrapolate=ext) for t in ts]
        store[f"unit/traj_pos_mode{mode}"] = np.array(
            [np.full(6, np.nan) if r is None else r for r in res]
        )
    store["unit/traj_vel"] = np.array([tr.velocity_at_t(float(t)) for t in ts])
    one = Trajectory(data[:1])
    store["unit/traj1_data"] = np.array(one.data)
    store["unit/traj1_pos_mode2"] = np.array([one.position_at_t(float(t), extrapolate=True) for t in ts])
    store["unit/traj1_pos_mode1"] = np.array([one.position_at_t(float(t)) for t in ts])
    # restated Shapely: box-pair intersects through the reference's detect_collisions
    from scenario_gym.state.utils import detect_collisions

    n = 400
    pa = np.zeros((n, 6))
    pb = np.zeros((n, 6))
    pa[:, 0:2] = rng.uniform(-4, 4, (n, 2))
    pb[:, 0:2] = rng.uniform(-4, 4, (n, 2))
    pa[:, 3] = rng.uniform(-np.pi, np.pi, n)
    pb[:, 3] = rng.uniform(-np.pi, np.pi, n)
    # exact touching / identical cases
    pa[:8] = 0.0
    pb[:8] = 0.0
    pb[0, 0] = 4.2  # edge to edge touching (closed set => collide)
    pb[1, 0] = 4.2 + 1e-12
    pb[2, 1] = 2.0  # side touching
    pb[3, 1] = np.nextafter(2.0, 3.0)
    pb[4, 0:2] = (4.2, 2.0)  # corner touching
    pb[5, 0:2] = (0.0, 0.0)  # identical boxes never collide (utils.py:58)
    pa[6, 3] = np.pi / 2
    pb[6, 0] = 3.1
    pa[7, 3] = np.pi / 4
    pb[7, 0] = 3.0
    e2 = Entity(ce, ref="y")
    hits = np.zeros(n, np.uint8)
    for i in range(n):
        c = detect_collisions({e: pa[i], e2: pb[i]})
        hits[i] = len(c[e]) > 0
    store["unit/pair_pose_a"] = pa
    store["unit/pair_pose_b"] = pb
    store["unit/pair_hit"] = hits
    manifest["unit"] = {"box": list(synthetic.CAR1_BOX), "pair_hits": int(hits.sum())}


# --------------------------------------------------------------------------- future collisions
def gen_future(store, manifest):
    """FutureCollisionDetector (sensor/common.py:60-105) on the test scenarios, every 20th tick."""
    from scenario_gym.sensor.common import FutureCollisionDetector

    files = sorted(glob.glob(os.path.join(SCEN_DIR, "*.xosc")))
    total = hits = 0
    for f in files:
        name = os.path.splitext(os.path.basename(f))[0]
        gym = ScenarioGym()
        gym.load_scenario(f, relabel=True)  # same inputs as xosc/<name>/in
        ego = gym.state.scenario.ego
        sensors = {h: FutureCollisionDetector(ego, horizon=h) for h in (5.0, 1.5)}
        ticks, times, flags = [], [], {h: [] for h in sensors}
        k = 0
        while True:
            if k % 20 == 0:
                ticks.append(k)
                times.append(gym.state.t)
                for h, sn in sensors.items():
                    flags[h].append(bool(sn._step(gym.state).future_collision))
            if gym.state.is_done:
                break
            gym.step()
            k += 1
        store[f"future/{name}/tick"] = np.array(ticks, np.int32)
        store[f"future/{name}/t"] = np.array(times, np.float64)
        for h in sensors:
            store[f"future/{name}/flag_h{h}"] = np.array(flags[h], np.uint8)
            total += len(flags[h])
            hits += int(np.sum(flags[h]))
        print("future", name, len(ticks), {h: int(np.sum(v)) for h, v in flags.items()})
    manifest["future"] = {"horizons": [5.0, 1.5], "n_samples": 10, "queries": total, "hits": hits}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])  # e.g. `python -m oracle.gen_golden pid` regenerates one file only
    mpath = os.path.join(GOLDEN, "manifest.json")
    manifest = json.load(open(mpath)) if (only and os.path.exists(mpath)) else {}
    manifest.update({"reference": "driskai/scenario_gym v0.3.1", "numpy": np.__version__})
    import scipy

    manifest["scipy"] = scipy.__version__

    def want(name):
        return not only or name in only

    if want("unit"):
        store = {}
        gen_unit(store, manifest)
        np.savez_compressed(os.path.join(GOLDEN, "unit.npz"), **store)
    if want("xosc"):
        store = {}
        gen_xosc(store, manifest)
        np.savez_compressed(os.path.join(GOLDEN, "xosc.npz"), **store)
    if want("veh_rss"):
        store = {}
        cfg = golden_cases.veh_cfg()
        gen_vehicle_like("veh", cfg, store, manifest)
        # terminal conditions "collision" / "ego_collision" (state/state.py:399-400)
        gen_vehicle_like("veh_term", cfg, store, manifest, terminal=["max_length", "collision"])
        gen_vehicle_like("veh_egoterm", cfg, store, manifest, terminal=["max_length", "ego_collision"])
        cfg = golden_cases.rss_cfg()
        gen_vehicle_like("rss", cfg, store, manifest, rss=True)
        np.savez_compressed(os.path.join(GOLDEN, "veh_rss.npz"), **store)
    if want("ped"):
        store = {}
        gen_ped(store, manifest)
        np.savez_compressed(os.path.join(GOLDEN, "ped.npz"), **store)
    if want("pid"):
        store = {}
        gen_pid(store, manifest)
        np.savez_compressed(os.path.join(GOLDEN, "pid.npz"), **store)
    if want("road"):
        store = {}
        gen_road(store, manifest)
        np.savez_compressed(os.path.join(GOLDEN, "road.npz"), **store)
    if want("future"):
        store = {}
        gen_future(store, manifest)
        np.savez_compressed(os.path.join(GOLDEN, "future.npz"), **store)

    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(GOLDEN)):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)))


if __name__ == "__main__":
    main()
