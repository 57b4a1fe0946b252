"""
sg_rollout_host (the end-to-end entry point: HOST buffers in, per-scenario results out) against the
resident path: chunked action-table upload, fp32 tables and the device-side action source must all
reproduce the resident rollout bit for bit.
"""
import numpy as np
import pytest

from scenario_gym_b200 import abi, synthetic
from scenario_gym_b200.synthetic import pack_synthetic

pytestmark = pytest.mark.gpu

FIELDS = ("ego_avg_speed", "ego_max_speed", "ego_dist", "first_coll_tick", "first_coll_pair", "n_pair_ticks",
          "rss_flags", "tick", "t")


def _params(dt, rss):
    p = abi.default_params()
    p.timestep = dt
    p.features = abi.FEAT_COLLISIONS | abi.FEAT_EGO_METRICS | (abi.FEAT_RSS if rss else 0)
    return p


@pytest.mark.parametrize("T", [5, 16, 40])  # below / at / above the 16-tick upload chunk
def test_host_path_equals_resident_path(T):
    from scenario_gym_b200.engine import Engine
    from scenario_gym_b200.hostpath import HostRollout

    cfg = synthetic.vehicles_config(seed=11, N=50, M=64, T=T, half_extent=70.0)
    scene = pack_synthetic(cfg)
    p = _params(cfg.dt, rss=True)
    ref = Engine(scene, p, device=0)
    ref.reset()
    ref.rollout(-1, actions=cfg.actions)
    want = {k: ref.get(k) for k in FIELDS}
    assert int(want["n_pair_ticks"].sum()) > 0
    for actions in (cfg.actions, cfg.action_rng):
        eng = Engine(scene, p, device=0)
        hr = HostRollout(eng, actions)
        for _ in range(2):  # a second call must start from a clean reset
            got = hr.run()
            for k in FIELDS:
                assert np.array_equal(got[k], want[k], equal_nan=True), (type(actions).__name__, k)
            assert int(got["event_count"][0]) == int(ref.tensor("event_count").item())
        for k in ("pose", "vel", "dist", "collided", "rss_state", "safe_dist"):
            assert np.array_equal(eng.get(k), ref.get(k), equal_nan=True), k
        assert hr.d2h_bytes > 0 and hr.h2d_bytes >= scene.traj_rows.nbytes + scene.box.nbytes
    # fp32 table == resident rollout of the widened table
    a32 = cfg.actions.astype(np.float32)
    ref.reset()
    ref.rollout(-1, actions=a32.astype(np.float64))
    eng = Engine(scene, p, device=0)
    got = HostRollout(eng, a32).run()
    for k in FIELDS:
        assert np.array_equal(got[k], ref.get(k), equal_nan=True), ("fp32", k)


def test_host_path_without_actions():
    """Replay-only and pedestrian scenes: one fused rollout after the scene upload."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import all_xosc_specs
    from scenario_gym_b200.engine import Engine
    from scenario_gym_b200.hostpath import HostRollout
    from scenario_gym_b200.packing import pack_scenarios

    scene = pack_scenarios([s for _, s, _, _ in all_xosc_specs("xosc")])
    p = abi.default_params()
    ref = Engine(scene, p, device=0)
    ref.reset()
    ref.rollout(-1)
    for device_union in (True, False):  # union table built on the device from the knot times / uploaded
        eng = Engine(scene, p, device=0)
        hr = HostRollout(eng, device_union=device_union)
        assert hr.device_union == device_union
        got = hr.run()
        for k in FIELDS:
            assert np.array_equal(got[k], ref.get(k), equal_nan=True), (device_union, k)
    assert HostRollout(eng, device_union=True).h2d_bytes < HostRollout(eng, device_union=False).h2d_bytes


def _union_scenes():
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import all_xosc_specs
    from scenario_gym_b200.packing import ScenarioSpec, SlotSpec, pack_scenarios

    scenes = [pack_scenarios([s for _, s, _, _ in all_xosc_specs("xosc")])]
    # single control points (duplicated 0.1 s later by the reference), knots shared between entities,
    # times before / after a trajectory's range
    rng = np.random.default_rng(3)
    specs = []
    for n in range(7):
        slots = []
        for s in range(5):
            K = [1, 2, 7, 30, 1][s] if n % 2 == 0 else int(rng.integers(1, 12))
            t = np.sort(rng.uniform(-1.0, 9.0, K)) if s % 2 else np.arange(K) * 0.5 + 0.25 * s
            tr = np.concatenate([t[:, None], rng.normal(0, 20, (K, 6))], axis=1)
            slots.append(SlotSpec(kind=abi.KIND_AGENT_REPLAY if s == 0 else abi.KIND_REPLAY, traj=tr))
        sp = ScenarioSpec(slots=slots)
        sp.finalize()
        specs.append(sp)
    scenes.append(pack_scenarios(specs))
    return scenes


def test_union_table_built_on_device():
    """sg_build_union_x == packing.build_union_table (entity/batch.py:80-112), bit for bit."""
    from scenario_gym_b200.engine import Engine

    for scene in _union_scenes():
        eng = Engine(scene, abi.default_params(), device=0)
        eng._scene_t["union_x"].fill_(-7.0)
        got = eng.build_union_on_device().cpu().numpy()
        assert got.shape == scene.union_x.shape and scene.union_x.size > 0
        assert np.array_equal(got, scene.union_x)


_STATE_FIELDS = ("coll_mask", "tick", "done", "present", "collided", "first_coll_tick", "first_coll_pair", "n_pair_ticks", "ego_hits",
                 "t", "prev_t", "pose", "vel", "dist", "speed", "ego_avg_speed", "ego_max_speed", "ego_dist", "rss_flags",
                 "rss_state", "rss_last", "safe_dist", "goal_idx", "force")


def _window_cases():
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import all_xosc_specs
    from scenario_gym_b200.packing import pack_scenarios

    cfg = synthetic.vehicles_config(seed=21, N=37, M=64, T=30, half_extent=60.0)
    yield "vehicles+rss (in-kernel actions)", pack_synthetic(cfg), _params(cfg.dt, rss=True), cfg.action_rng
    cfg = synthetic.highway_config(seed=22, N=9, M=256, T=12)
    cfg.x0[:] = cfg.x0 * 0.5
    yield "sorted sweep", pack_synthetic(cfg), _params(cfg.dt, rss=True), cfg.action_rng
    yield "replay", pack_scenarios([s for _, s, _, _ in all_xosc_specs("xosc")]), abi.default_params(), None
    cfg = synthetic.crowd_config(seed=23, N=5, M=320, T=8, side=14.0)
    yield "crowd", pack_synthetic(cfg), _params(cfg.dt, rss=False), None
    cfg = synthetic.crowd_config(seed=24, N=6, M=96, T=10, side=9.0)
    yield "pedestrians (general kernel)", pack_synthetic(cfg), _params(cfg.dt, rss=False), None
    # per-agent VehicleController limits route the scene to the general kernel (device-side action source,
    # RSS, per-slot limit planes and the pair matrix all addressed through the window)
    cfg = synthetic.vehicles_config(seed=25, N=11, M=40, T=20, half_extent=35.0)
    scene = pack_synthetic(cfg)
    lim = np.zeros((4, scene.N * scene.M))
    rng = np.random.default_rng(7)
    lim[0], lim[1], lim[2], lim[3] = rng.uniform(0.3, 0.9, lim.shape[1]), rng.uniform(2.0, 6.0, lim.shape[1]), np.nan, 1.0
    lim[2, ::3] = 12.0
    scene.veh_limits = lim
    p = _params(cfg.dt, rss=True)
    p.features |= abi.FEAT_COLL_MATRIX
    yield "vehicles with per-agent limits + pair matrix (general kernel)", scene, p, cfg.action_rng


@pytest.mark.parametrize("windows", ["2", "3", "4", "5", "6"])
def test_host_path_scenario_windows(windows, monkeypatch):
    """
    sg_rollout_host uploads a batch in windows of scenarios and rolls every window out as it arrives
    (SgScene.plane_stride / scenario_base): every State row, the action stream's draws and the event
    records must equal the one-piece rollout bit for bit.  (Replay-only batches of three or more windows
    take the upload-bound policy: even windows, per-slot / per-scenario arrays sent once with the first.)
    """
    from scenario_gym_b200.engine import Engine
    from scenario_gym_b200.hostpath import HostRollout

    for tag, scene, p, actions in _window_cases():
        ref = Engine(scene, p, device=0)
        monkeypatch.setenv("SG_HOST_WINDOWS", "1")
        want = HostRollout(ref, actions).run()
        eng = Engine(scene, p, device=0)
        monkeypatch.setenv("SG_HOST_WINDOWS", windows)
        got = HostRollout(eng, actions).run()
        for k in FIELDS:
            assert np.array_equal(got[k], want[k], equal_nan=True), (tag, k)
        assert int(got["event_count"][0]) == int(want["event_count"][0]), tag
        for k in _STATE_FIELDS:
            assert np.array_equal(eng.get(k), ref.get(k), equal_nan=True), (tag, k)
        a, b = eng.events(), ref.events()
        assert a.tobytes() == b.tobytes(), (tag, "events")


def test_host_path_default_window_policy_large_replay_batch(monkeypatch):
    """
    A replay-only batch above the windowing threshold (8 MiB, 64 scenarios) with NO override takes the
    upload-bound policy (five even windows, the per-slot / per-scenario arrays of the whole batch sent once
    with the first): results equal the one-piece upload's bit for bit.
    """
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import all_xosc_specs
    from scenario_gym_b200.engine import Engine
    from scenario_gym_b200.hostpath import HostRollout
    from scenario_gym_b200.packing import pack_scenarios, tile_scene

    scene = tile_scene(pack_scenarios([s for _, s, _, _ in all_xosc_specs("xosc")]), 40)
    assert scene.N >= 64 and scene.traj_rows.nbytes >= (8 << 20)
    p = abi.default_params()
    ref = Engine(scene, p, device=0)
    monkeypatch.setenv("SG_HOST_WINDOWS", "1")
    want = HostRollout(ref).run()
    want = {k: v.copy() for k, v in want.items()}
    monkeypatch.delenv("SG_HOST_WINDOWS")
    eng = Engine(scene, p, device=0)
    got = HostRollout(eng).run()
    for k in FIELDS:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    for k in _STATE_FIELDS:
        assert np.array_equal(eng.get(k), ref.get(k), equal_nan=True), k
    assert eng.events().tobytes() == ref.events().tobytes()


def test_host_path_default_window_policy_large_vehicle_batch(monkeypatch):
    """
    A compute-bound batch of 4096+ scenarios with NO override goes up in four windows (1/32, 1/8, 1/2, 1 of the
    batch; in-kernel action stream addressed through the windows): equals the one-piece upload bit for bit.
    """
    from scenario_gym_b200.engine import Engine
    from scenario_gym_b200.hostpath import HostRollout

    cfg = synthetic.vehicles_config(seed=31, N=4100, M=64, T=10, half_extent=60.0, materialise=False)
    scene = pack_synthetic(cfg)
    assert scene.traj_rows.nbytes >= (8 << 20)
    p = _params(cfg.dt, rss=True)
    ref = Engine(scene, p, device=0)
    monkeypatch.setenv("SG_HOST_WINDOWS", "1")
    want = {k: v.copy() for k, v in HostRollout(ref, cfg.action_rng).run().items()}
    monkeypatch.delenv("SG_HOST_WINDOWS")
    eng = Engine(scene, p, device=0)
    got = HostRollout(eng, cfg.action_rng).run()
    for k in FIELDS:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    for k in ("tick", "done", "collided", "pose", "vel", "dist", "rss_state", "rss_last", "safe_dist"):
        assert np.array_equal(eng.get(k), ref.get(k), equal_nan=True), k
    assert eng.events().tobytes() == ref.events().tobytes()
    assert int(got["n_pair_ticks"].sum()) > 0
