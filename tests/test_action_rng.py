"""
Device-side action source (SgActionRng): the PCG64 restatements (oracle C, CUDA) against numpy's own
generator, and rollouts that draw their actions in-kernel against rollouts fed the numpy table.
"""
import numpy as np
import pytest

from oracle.runner import OracleEngine
from scenario_gym_b200 import abi, synthetic
from scenario_gym_b200.action_rng import ActionRng
from scenario_gym_b200.packing import ScenarioSpec, SlotSpec, pack_scenarios
from scenario_gym_b200.synthetic import pack_synthetic

STATE_KEYS = ("tick", "done", "present", "collided", "first_coll_tick", "first_coll_pair", "n_pair_ticks",
              "ego_hits", "t", "prev_t", "pose", "vel", "dist", "speed", "ego_avg_speed", "ego_max_speed",
              "ego_dist", "rss_flags", "rss_state", "rss_last", "safe_dist", "safe_ratio")


def _rng_cases():
    big = ActionRng.from_generator(12345, offset=(3 * 10 ** 12 + 7, 5 * 10 ** 13), tick_stride=2 ** 33 + 1,
                                   low=(-6.0, -1.0), high=(6.0, 1.0), n_ticks=5, nm=37)
    small = ActionRng.from_generator(np.random.default_rng(0), offset=(0, 6 * 50), tick_stride=50,
                                     low=(-2.0, -0.02), high=(2.0, 0.02), n_ticks=6, nm=50)
    return [big, small, small.shard(13, 20)]


def test_numpy_uniform_is_low_plus_range_times_random():
    """The two-rounding form low + (high - low) * u equals Generator.uniform bit for bit."""
    for lo, hi in ((-6.0, 6.0), (-1.0, 1.0), (-2.0, 2.0), (-0.02, 0.02), (0.1, 0.7)):
        a = np.random.default_rng(5).uniform(lo, hi, 1 << 20)
        u = np.random.default_rng(5).random(1 << 20)
        assert np.array_equal(a, lo + (hi - lo) * u)
        tmp = u.copy()
        tmp *= hi - lo
        tmp += lo
        assert np.array_equal(a, tmp)


def test_oracle_pcg64_matches_numpy(oracle_lib):
    import ctypes as C

    for r in _rng_cases():
        want = r.table()
        got = np.empty_like(want)
        st = r.struct()
        assert oracle_lib["fill_random_actions"](C.byref(st), 0, r.n_ticks, r.nm, got.ctypes.data, 0, None) == 0
        assert np.array_equal(got, want)
        part = np.empty((2, 2, r.nm))
        assert oracle_lib["fill_random_actions"](C.byref(st), 3, 2, r.nm, part.ctypes.data, 0, None) == 0
        assert np.array_equal(part, want[3:5])


def _veh_params(cfg, rss=True, matrix=False):
    p = abi.default_params()
    p.timestep = cfg.dt
    p.features = abi.FEAT_COLLISIONS | abi.FEAT_EGO_METRICS | (abi.FEAT_RSS if rss else 0) | \
        (abi.FEAT_COLL_MATRIX if matrix else 0)
    return p


def _same_state(a, b, what):
    for k in STATE_KEYS:
        assert np.array_equal(a.get(k), b.get(k), equal_nan=True), f"{what}: {k} differs"
    ea, eb = a.events(), b.events()
    assert np.array_equal(ea, eb), f"{what}: events differ"


def test_oracle_rollout_rng_equals_table():
    cfg = synthetic.highway_config(seed=2, N=6, M=16, T=20)
    cfg.x0[:] *= 0.4
    scene = pack_synthetic(cfg)
    engs = []
    for actions in (cfg.actions, cfg.action_rng, cfg.actions.astype(np.float32).astype(np.float64),
                    cfg.actions.astype(np.float32)):
        e = OracleEngine(scene, _veh_params(cfg))
        e.reset()
        e.rollout(-1, actions=actions)
        engs.append(e)
    _same_state(engs[0], engs[1], "oracle rng vs table")
    _same_state(engs[2], engs[3], "oracle f32 table vs widened table")
    assert int(engs[0].get("tick").min()) == cfg.T


def test_action_rows_bound_only_vehicle_scenarios():
    """A short action table stops the scenarios that consume it, not the replay-only ones next to them."""
    T = 12
    t_end = 3.05
    line = np.array([[0.0, 0.0, 0.0, 0, 0, 0, 0], [t_end, 30.0, 0.0, 0, 0, 0, 0]])
    other = line + np.array([0, 0, 50.0, 0, 0, 0, 0])
    veh = ScenarioSpec(slots=[SlotSpec(kind=abi.KIND_VEHICLE, traj=line), SlotSpec(kind=abi.KIND_REPLAY, traj=other)])
    rep = ScenarioSpec(slots=[SlotSpec(kind=abi.KIND_AGENT_REPLAY, traj=line), SlotSpec(kind=abi.KIND_REPLAY, traj=other)])
    scene = pack_scenarios([veh, rep])
    p = abi.default_params()
    p.timestep = 0.1
    e = OracleEngine(scene, p)
    e.reset()
    e.rollout(-1, actions=np.zeros((T, 2, scene.N * scene.M)))
    assert e.get("tick").tolist() == [T, 30]
    assert e.get("done").tolist() == [0, 1]


# ------------------------------------------------------------------------------------- GPU
def _gpu(scene, p, **kw):
    from scenario_gym_b200.engine import Engine

    return Engine(scene, p, device=0, **kw)


@pytest.mark.gpu
def test_gpu_pcg64_matches_numpy():
    cfg = synthetic.vehicles_config(seed=0, N=2, M=4, T=2)
    eng = _gpu(pack_synthetic(cfg), _veh_params(cfg))
    for r in _rng_cases():
        eng.N, eng.M = 1, r.nm  # fill_actions only checks nm against the engine's slot count
        want = r.table()
        assert np.array_equal(eng.fill_actions(r).cpu().numpy(), want)
        assert np.array_equal(eng.fill_actions(r, 3, 2).cpu().numpy(), want[3:5])


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["c3", "c3_norss", "c5", "subwarp", "m512"])
def test_gpu_rollout_rng_equals_table(shape):
    rss = shape != "c3_norss"
    if shape.startswith("c3"):
        cfg = synthetic.vehicles_config(seed=4, N=40, M=64, T=48, half_extent=60.0)
    elif shape == "c5":
        cfg = synthetic.highway_config(seed=4, N=9, M=256, T=40)
        cfg.x0[:] *= 0.5
    elif shape == "m512":
        cfg = synthetic.vehicles_config(seed=4, N=3, M=512, T=12, half_extent=100.0)
    else:
        cfg = synthetic.vehicles_config(seed=4, N=33, M=8, T=48, half_extent=15.0)
    scene = pack_synthetic(cfg)
    p = _veh_params(cfg, rss=rss)
    tab = _gpu(scene, p)
    tab.reset()
    tab.rollout(-1, actions=cfg.actions)
    assert int(tab.get("n_pair_ticks").sum()) > 0, "the case should exercise collisions"
    # fused, in-kernel stream
    rng = _gpu(scene, p)
    rng.reset()
    rng.rollout(-1, actions=cfg.action_rng)
    _same_state(tab, rng, f"{shape}: fused rng vs table")
    # partial rollouts: rows continue at tick0
    part = _gpu(scene, p)
    part.reset()
    k = 0
    for n in (1, 5, 2, cfg.T):
        n = min(n, cfg.T - k)
        part.rollout(n, actions=cfg.action_rng, tick0=k)
        k += n
    _same_state(tab, part, f"{shape}: chunked rng vs table")
    # fp32 tables: equal to the table of the widened values
    a32 = cfg.actions.astype(np.float32)
    w = _gpu(scene, p)
    w.reset()
    w.rollout(-1, actions=a32.astype(np.float64))
    f = _gpu(scene, p)
    f.reset()
    f.rollout(-1, actions=a32)
    _same_state(w, f, f"{shape}: fp32 table vs widened")
    # and against the CPU oracle consuming numpy's table
    cpu = OracleEngine(scene, p, event_cap=1 << 16)
    cpu.reset()
    cpu.rollout(-1, actions=cfg.actions)
    for k_ in ("tick", "collided", "first_coll_tick", "first_coll_pair", "n_pair_ticks", "rss_flags"):
        assert np.array_equal(rng.get(k_), cpu.get(k_)), k_
    assert np.allclose(rng.get("pose"), cpu.get("pose"), rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_gpu_rng_with_trace_and_in_mixed_scenes():
    """Non-lean vehicle rollouts (trace) materialise the rows; mixed scenes draw them in the general kernel."""
    cfg = synthetic.vehicles_config(seed=7, N=5, M=16, T=20, half_extent=20.0)
    scene = pack_synthetic(cfg)
    p = _veh_params(cfg, rss=False)
    a = _gpu(scene, p, trace_cap=cfg.T + 2)
    a.reset()
    a.rollout(-1, actions=cfg.actions)
    b = _gpu(scene, p, trace_cap=cfg.T + 2)
    b.reset()
    b.rollout(-1, actions=cfg.action_rng)
    _same_state(a, b, "trace rollout: rng vs table")
    assert np.array_equal(a.get("trace_pose"), b.get("trace_pose"))
    # mixed scene: make slot 3 of every scenario a replayed agent -> general kernel
    cfg.kind[:, 3] = abi.KIND_AGENT_REPLAY
    scene = pack_synthetic(cfg)
    outs = []
    for actions in (cfg.actions, cfg.action_rng, cfg.actions.astype(np.float32)):
        e = _gpu(scene, p)
        e.reset()
        e.rollout(-1, actions=actions)
        outs.append(e)
    _same_state(outs[0], outs[1], "mixed scene: rng vs table")
    cpu = OracleEngine(scene, p, event_cap=1 << 16)
    cpu.reset()
    cpu.rollout(-1, actions=cfg.actions.astype(np.float32))
    for k_ in ("tick", "collided", "first_coll_tick", "n_pair_ticks"):
        assert np.array_equal(outs[2].get(k_), cpu.get(k_)), k_
    assert np.allclose(outs[2].get("pose"), cpu.get("pose"), rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_gpu_action_rows_bound_only_vehicle_scenarios():
    T = 12
    t_end = 3.05
    line = np.array([[0.0, 0.0, 0.0, 0, 0, 0, 0], [t_end, 30.0, 0.0, 0, 0, 0, 0]])
    other = line + np.array([0, 0, 50.0, 0, 0, 0, 0])
    veh = ScenarioSpec(slots=[SlotSpec(kind=abi.KIND_VEHICLE, traj=line), SlotSpec(kind=abi.KIND_REPLAY, traj=other)])
    rep = ScenarioSpec(slots=[SlotSpec(kind=abi.KIND_AGENT_REPLAY, traj=line), SlotSpec(kind=abi.KIND_REPLAY, traj=other)])
    scene = pack_scenarios([veh, rep])
    p = abi.default_params()
    p.timestep = 0.1
    e = _gpu(scene, p)
    e.reset()
    e.rollout(-1, actions=np.zeros((T, 2, scene.N * scene.M)))
    assert e.get("tick").tolist() == [T, 30]
    assert e.get("done").tolist() == [0, 1]
