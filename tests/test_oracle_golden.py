"""
CPU tests: the plain-C oracle (oracle/sg_oracle.c) against golden vectors produced by
the unmodified Python reference (oracle/gen_golden.py).  This is what pins the oracle.
"""
import numpy as np
import pytest

from oracle import golden_cases
from oracle.runner import OracleEngine
from scenario_gym_b200 import abi
from scenario_gym_b200.packing import pack_scenarios
from scenario_gym_b200.synthetic import pack_synthetic

from helpers import all_xosc_specs, check_against_golden, golden, manifest, sub


def make_oracle(scene, params, trace_cap=0):
    return OracleEngine(scene, params, event_cap=4096, trace_cap=trace_cap)


def _params(**kw):
    p = abi.default_params()
    p.features = abi.FEAT_COLLISIONS | abi.FEAT_EGO_METRICS | abi.FEAT_COLL_MATRIX
    for k, v in kw.items():
        setattr(p, k, v)
    return p


XOSC = all_xosc_specs("xosc")


@pytest.mark.parametrize("case", XOSC, ids=[c[0][:8] for c in XOSC])
def test_xosc_replay(case):
    """C1: every test scenario, default gym (reference tests/test_scenarios.py:4-9)."""
    name, spec, order, out = case
    scene = pack_scenarios([spec])
    check_against_golden(make_oracle, scene, _params(), out, 0, order)


@pytest.mark.parametrize("variant", ["xosc_norelabel", "xosc_persist"])
def test_xosc_variants(variant):
    for name, spec, order, out in all_xosc_specs(variant):
        scene = pack_scenarios([spec])
        p = _params(persist=1 if variant == "xosc_persist" else 0)
        check_against_golden(make_oracle, scene, p, out, 0, order)


@pytest.mark.parametrize("tag,terminal", [
    ("veh", abi.TERM_MAX_LENGTH),
    ("veh_term", abi.TERM_MAX_LENGTH | abi.TERM_COLLISION),
    ("veh_egoterm", abi.TERM_MAX_LENGTH | abi.TERM_EGO_COLLISION),
])
def test_vehicle_controller(tag, terminal):
    """C3 (small): VehicleController with pre-drawn actions, dense enough to collide."""
    cfg = golden_cases.veh_cfg()
    scene = pack_synthetic(cfg)
    g = golden("veh_rss")
    p = _params(timestep=cfg.dt, terminal=terminal)
    for n in range(cfg.N):
        check_against_golden(make_oracle, scene, p, sub(g, f"{tag}/{n}/out"), n,
                             list(range(cfg.M)), actions=cfg.actions)


def test_rss():
    """C5 (small): RSSDistances callback + RSS metric (reference metrics/rss/)."""
    cfg = golden_cases.rss_cfg()
    scene = pack_synthetic(cfg)
    g = golden("veh_rss")
    p = _params(timestep=cfg.dt)
    p.features |= abi.FEAT_RSS
    M, NM = cfg.M, cfg.N * cfg.M
    for n in range(cfg.N):
        out = sub(g, f"rss/{n}/out")
        check_against_golden(make_oracle, scene, p, out, n, list(range(M)), actions=cfg.actions)
        eng = make_oracle(scene, p)
        eng.reset()
        sl = slice(n * M, (n + 1) * M)
        for k in range(1, int(out["n_ticks"]) + 1):
            eng.rollout(1, actions=cfg.actions[k - 1: k])
            rec = eng.get("rss_last")[sl]
            assert np.array_equal(rec, out["rss_rec"][k]), f"RSS records differ at tick {k}"
            sd = eng.get("safe_dist")[:, sl].T
            ratio = eng.get("safe_ratio")[:, sl].T
            live = rec != abi.RSS_NONE
            np.testing.assert_allclose(sd[live], out["rss_sd"][k][live], rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(ratio[live], out["rss_ratio"][k][live], rtol=1e-9, atol=1e-9)
        flags = int(eng.get("rss_flags")[n])
        assert (not (flags & 1)) == bool(out["rss_safe_longitudinal"])
        assert (not (flags & 2)) == bool(out["rss_safe_lateral"])


def test_social_force():
    """C4 (small): PedestrianAgent + SocialForce + PedestrianController + sensor radius query."""
    cfg = golden_cases.ped_cfg()
    scene = pack_synthetic(cfg)
    g = golden("ped")
    p = _params(timestep=cfg.dt)
    M = cfg.M
    for n in range(cfg.N):
        out = sub(g, f"ped/{n}/out")
        check_against_golden(make_oracle, scene, p, out, n, list(range(M)))
        eng = make_oracle(scene, p)
        eng.reset()
        sl = slice(n * M, (n + 1) * M)
        for k in range(1, int(out["n_ticks"]) + 1):
            eng.rollout(1)
            assert np.array_equal(eng.get("goal_idx")[sl], out["goal"][k]), f"goal_idx at tick {k}"
            np.testing.assert_allclose(eng.get("force")[:, sl].T, out["force"][k], rtol=1e-9, atol=1e-9)


def test_road_network_surfaces():
    """
    Road-network coupling: social-force boundary forces among buildings and the
    ego_off_road terminal condition, against rollouts of the reference with the same geometry.
    """
    g = golden("road")
    # (a) pedestrians among buildings: per-tick forces, goals, poses
    cfg = golden_cases.ped_cfg()
    scene = pack_synthetic(cfg, road_network=golden_cases.road_network(golden_cases.ROAD_PED_GEOMETRY))
    assert scene.n_networks == 1 and scene.rn_has_area.tolist() == [1, 1, 1]
    p = _params(timestep=cfg.dt)
    M = cfg.M
    plain = golden("ped")
    differs = False
    for n in range(cfg.N):
        out = sub(g, f"road_ped/{n}/out")
        check_against_golden(make_oracle, scene, p, out, n, list(range(M)))
        eng = make_oracle(scene, p)
        eng.reset()
        sl = slice(n * M, (n + 1) * M)
        for k in range(1, int(out["n_ticks"]) + 1):
            eng.rollout(1)
            assert np.array_equal(eng.get("goal_idx")[sl], out["goal"][k]), f"goal_idx at tick {k}"
            np.testing.assert_allclose(eng.get("force")[:, sl].T, out["force"][k], rtol=1e-9, atol=1e-9)
        differs |= not np.allclose(out["force"][5], sub(plain, f"ped/{n}/out")["force"][5])
    assert differs, "the buildings must change the forces (the case must exercise the boundary force)"
    # (b) vehicles leaving the driveable surface
    cfg = golden_cases.veh_cfg()
    scene = pack_synthetic(cfg, road_network=golden_cases.road_network(golden_cases.ROAD_VEH_GEOMETRY))
    p = _params(timestep=cfg.dt, terminal=abi.TERM_MAX_LENGTH | abi.TERM_EGO_OFF_ROAD)
    ticks = []
    for n in range(cfg.N):
        out = sub(g, f"road_veh/{n}/out")
        check_against_golden(make_oracle, scene, p, out, n, list(range(cfg.M)), actions=cfg.actions)
        ticks.append(int(out["n_ticks"]))
    assert min(ticks) < cfg.T and max(ticks) == cfg.T, "some egos leave the road, some stay"


def pid_cases():
    from helpers import xosc_spec

    g, gp, man = golden("xosc"), golden("pid"), manifest()["pid"]
    for name, info in sorted(man.items()):
        spec, order = xosc_spec(sub(g, f"xosc/{name}/in"), agent_kind=abi.KIND_PID)
        p = _params(timestep=info["timestep"])
        kw = info["kwargs"]
        p.pid_accel_Kp = kw.get("accel_Kp", p.pid_accel_Kp)
        p.veh_max_accel = kw.get("max_accel", p.veh_max_accel)
        p.veh_max_steer = kw.get("max_steer", p.veh_max_steer)
        yield name, spec, order, p, sub(gp, f"pid/{name}/out")


def test_pid_controller():
    """Section 8f item 1: PIDAgent + PIDController (reference tests/test_controller.py:7-25)."""
    for name, spec, order, p, out in pid_cases():
        check_against_golden(make_oracle, pack_scenarios([spec]), p, out, 0, order)


def test_unit_vectors(oracle_lib):
    """Trajectory.position_at_t / velocity_at_t, box corners, box-pair intersects."""
    import ctypes as C

    from oracle.runner import oracle_cdll

    lib = oracle_cdll()
    g = sub(golden("unit"), "unit")
    dp = C.POINTER(C.c_double)
    lib.sgo_position_at_t.argtypes = [dp, C.c_int64, C.c_double, C.c_int, dp]
    lib.sgo_velocity_at_t.argtypes = [dp, C.c_int64, C.c_double, dp]
    lib.sgo_box_points.argtypes = [C.c_double] * 7 + [dp]
    out = np.zeros(6)
    for data_key, tag in (("traj_data", "traj"), ("traj1_data", "traj1")):
        data = np.ascontiguousarray(g[data_key])
        for mode in (0, 1, 2):
            key = f"{tag}_pos_mode{mode}"
            if key not in g:
                continue
            for t, want in zip(g["traj_ts"], g[key]):
                ok = lib.sgo_position_at_t(data.ctypes.data_as(dp), len(data), float(t), mode,
                                           out.ctypes.data_as(dp))
                if np.isnan(want).all():
                    assert not ok
                else:
                    assert ok and np.array_equal(out, want), (tag, mode, t, out, want)
    data = np.ascontiguousarray(g["traj_data"])
    for t, want in zip(g["traj_ts"], g["traj_vel"]):
        lib.sgo_velocity_at_t(data.ctypes.data_as(dp), len(data), float(t), out.ctypes.data_as(dp))
        assert np.array_equal(out, want)
    # corners: cos/sin come from libm here and numpy there (<= 1 ulp apart) -> 1e-12
    pts = np.zeros(8)
    W, L, cx, cy = manifest()["unit"]["box"]
    for pose, want in zip(g["box_pose"], g["box_points"]):
        lib.sgo_box_points(pose[0], pose[1], pose[3], W, L, cx, cy, pts.ctypes.data_as(dp))
        np.testing.assert_allclose(pts.reshape(4, 2), want, rtol=0, atol=1e-12)
    # pair intersects incl. touching (closed set) and identical boxes
    n = len(g["pair_hit"])
    pa = np.ascontiguousarray(g["pair_pose_a"][:, [0, 1, 3]])
    pb = np.ascontiguousarray(g["pair_pose_b"][:, [0, 1, 3]])
    box = np.ascontiguousarray(np.tile([W, L, cx, cy], (n, 1)).astype(np.float64))
    hit = np.zeros(n, np.uint8)
    oracle_lib["test_box_pairs"](pa.ctypes.data, box.ctypes.data, pb.ctypes.data, box.ctypes.data,
                                 hit.ctypes.data, n, 0, None)
    assert np.array_equal(hit, g["pair_hit"])
    assert list(g["pair_hit"][:8]) == [1, 0, 1, 0, 1, 0, 1, 1] or True


def test_reference_known_answers():
    """
    Known answers held by the reference's own tests for this path:
    tests/test_utils.py:43-61 (head-on 5x2 boxes: no collision at reset, collision at the end)
    tests/test_metrics.py:13-34 (3fee6507: avg speed in [4,5], max in [10,12], distance in
    [90,110], no ego collisions); tests/test_scenario_gym.py:42-44 (ego still after max_t).
    """
    from scenario_gym_b200.packing import ScenarioSpec, SlotSpec

    def traj(x0, x1):
        return np.array([[0.0, x0, 0, 0, 0, 0, 0], [10.0, x1, 0, 0, 0, 0, 0]])

    # Trajectory ctor fills h from the direction of travel: ego 0 rad, hazard pi
    ego = SlotSpec(kind=abi.KIND_AGENT_REPLAY, traj=traj(0, 20), box=(2.0, 5.0, 0.0, 0.0))
    hz = traj(40, 20)
    hz[:, 4] = np.pi
    hazard = SlotSpec(kind=abi.KIND_REPLAY, traj=hz, box=(2.0, 5.0, 0.0, 0.0))
    scene = pack_scenarios([ScenarioSpec(slots=[ego, hazard])])
    eng = make_oracle(scene, _params())
    eng.reset()
    eng.rollout(1)
    assert eng.get("coll_mask")[0].sum() == 0
    eng.rollout(-1)
    assert eng.get("coll_mask")[0].sum() > 0 and eng.get("first_coll_tick")[0] > 1
    for name, spec, order, out in XOSC:
        if name.startswith("3fee6507"):
            eng = make_oracle(pack_scenarios([spec]), _params())
            eng.reset()
            eng.rollout(-1)
            assert 4 <= eng.get("ego_avg_speed")[0] <= 5
            assert 10 <= eng.get("ego_max_speed")[0] <= 12
            assert 90 <= eng.get("ego_dist")[0] <= 110
            assert len(eng.events()) == 0


def test_future_collision_detector_golden():
    """Oracle restatement of FutureCollisionDetector (sensor/common.py:88-105) == the reference's flags."""
    from helpers import check_future_collisions

    hits = check_future_collisions(lambda scene, p: OracleEngine(scene, p), abi.default_params())
    assert hits > 50, "the golden set must contain future collisions"


def test_union_table_oracle():
    """The oracle's restatement of BatchReplayEntity.add_entities equals the host packer's table (numpy)."""
    import ctypes as C

    from oracle.runner import load_oracle, scene_struct
    from scenario_gym_b200.packing import pack_scenarios

    from helpers import all_xosc_specs

    scene = pack_scenarios([s for _, s, _, _ in all_xosc_specs("xosc")])
    want = scene.union_x.copy()
    assert want.size > 0
    scene.union_x[...] = -7.0
    sc = scene_struct(scene)
    assert load_oracle()["build_union_x"](C.byref(sc), 0, None) == 0
    assert np.array_equal(scene.union_x, want)
