"""
GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI, against
(a) the reference's golden vectors and (b) the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): collision flags, first-collision tick, pairs and events
bit-exact; poses and continuous metrics within 1e-9 absolute-or-relative in fp64.
"""
import numpy as np
import pytest

from oracle import golden_cases
from oracle.runner import OracleEngine
from scenario_gym_b200 import abi, synthetic
from scenario_gym_b200.packing import pack_scenarios, tile_scene
from scenario_gym_b200.synthetic import pack_synthetic

from helpers import all_xosc_specs, check_against_golden, check_future_collisions, golden, manifest, sub

pytestmark = pytest.mark.gpu

TOL = 1e-9


def make_gpu(scene, params, trace_cap=0):
    from scenario_gym_b200.engine import Engine

    return Engine(scene, params, device=0, event_cap=1 << 16, trace_cap=trace_cap)


def _params(**kw):
    p = abi.default_params()
    p.features = abi.FEAT_COLLISIONS | abi.FEAT_EGO_METRICS | abi.FEAT_COLL_MATRIX
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def close(a, b, what, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    err[np.isnan(a) & np.isnan(b)] = 0.0
    err[(a == b)] = 0.0
    assert np.all(err <= tol), f"{what}: max err {np.nanmax(err):.3e}"


def compare_engines(gpu, cpu, scene, what, check_rss=False, check_ped=False):
    """Final state of a GPU engine vs the CPU oracle after identical calls."""
    for k in ("tick", "done", "present", "collided", "first_coll_tick", "first_coll_pair",
              "n_pair_ticks", "ego_hits"):
        assert np.array_equal(gpu.get(k), cpu.get(k)), f"{what}: {k} differs"
    assert np.array_equal(gpu.get("t"), cpu.get("t")), f"{what}: tick times differ"
    ge, ce = gpu.events(), cpu.events()
    assert len(ge) == len(ce), f"{what}: {len(ge)} vs {len(ce)} events"
    for f in ("scenario", "tick", "slot", "t"):
        assert np.array_equal(ge[f], ce[f]), f"{what}: event {f} differs"
    pres = cpu.get("present").astype(bool)
    for k in ("pose", "vel"):
        close(gpu.get(k)[:, pres], cpu.get(k)[:, pres], f"{what}: {k}")
    for k in ("dist", "speed", "ego_avg_speed", "ego_max_speed", "ego_dist"):
        close(gpu.get(k), cpu.get(k), f"{what}: {k}")
    if check_rss:
        assert np.array_equal(gpu.get("rss_flags"), cpu.get("rss_flags")), f"{what}: rss flags"
        assert np.array_equal(gpu.get("rss_state"), cpu.get("rss_state")), f"{what}: rss state"
        assert np.array_equal(gpu.get("rss_last"), cpu.get("rss_last")), f"{what}: rss record"
        fin = np.isfinite(cpu.get("safe_ratio"))
        close(gpu.get("safe_dist"), cpu.get("safe_dist"), f"{what}: safe_dist")
        close(gpu.get("safe_ratio")[fin], cpu.get("safe_ratio")[fin], f"{what}: safe_ratio")
    if check_ped:
        assert np.array_equal(gpu.get("goal_idx"), cpu.get("goal_idx")), f"{what}: goal_idx"
        close(gpu.get("force"), cpu.get("force"), f"{what}: force")


# ----------------------------------------------------------------------------- golden vectors
XOSC = all_xosc_specs("xosc")


@pytest.mark.parametrize("case", XOSC, ids=[c[0][:8] for c in XOSC])
def test_xosc_replay_golden(case):
    name, spec, order, out = case
    check_against_golden(make_gpu, pack_scenarios([spec]), _params(), out, 0, order)


@pytest.mark.parametrize("variant", ["xosc_norelabel", "xosc_persist"])
def test_xosc_variants_golden(variant):
    for name, spec, order, out in all_xosc_specs(variant):
        p = _params(persist=1 if variant == "xosc_persist" else 0)
        check_against_golden(make_gpu, pack_scenarios([spec]), p, out, 0, order)


@pytest.mark.parametrize("tag,terminal", [
    ("veh", abi.TERM_MAX_LENGTH),
    ("veh_term", abi.TERM_MAX_LENGTH | abi.TERM_COLLISION),
    ("veh_egoterm", abi.TERM_MAX_LENGTH | abi.TERM_EGO_COLLISION),
])
def test_vehicle_golden(tag, terminal):
    cfg = golden_cases.veh_cfg()
    scene = pack_synthetic(cfg)
    g = golden("veh_rss")
    p = _params(timestep=cfg.dt, terminal=terminal)
    for n in range(cfg.N):
        check_against_golden(make_gpu, scene, p, sub(g, f"{tag}/{n}/out"), n, list(range(cfg.M)),
                             actions=cfg.actions)


def test_rss_golden():
    cfg = golden_cases.rss_cfg()
    scene = pack_synthetic(cfg)
    g = golden("veh_rss")
    p = _params(timestep=cfg.dt)
    p.features |= abi.FEAT_RSS
    M = cfg.M
    eng = make_gpu(scene, p)
    eng.reset()
    T = cfg.T
    for k in range(1, T + 1):
        eng.rollout(1, actions=cfg.actions[k - 1: k])
        rec = eng.get("rss_last").reshape(cfg.N, M)
        sd = eng.get("safe_dist").reshape(2, cfg.N, M)
        ratio = eng.get("safe_ratio").reshape(2, cfg.N, M)
        for n in range(cfg.N):
            out = sub(g, f"rss/{n}/out")
            assert np.array_equal(rec[n], out["rss_rec"][k]), f"RSS records differ, scenario {n} tick {k}"
            live = rec[n] != abi.RSS_NONE
            close(sd[:, n].T[live], out["rss_sd"][k][live], "safe distances")
            close(ratio[:, n].T[live], out["rss_ratio"][k][live], "safe ratios")
    flags = eng.get("rss_flags")
    for n in range(cfg.N):
        out = sub(g, f"rss/{n}/out")
        assert (not (flags[n] & 1)) == bool(out["rss_safe_longitudinal"])
        assert (not (flags[n] & 2)) == bool(out["rss_safe_lateral"])
        check_against_golden(make_gpu, scene, p, out, n, list(range(M)), actions=cfg.actions)


def test_social_force_golden():
    cfg = golden_cases.ped_cfg()
    scene = pack_synthetic(cfg)
    g = golden("ped")
    p = _params(timestep=cfg.dt)
    M = cfg.M
    eng = make_gpu(scene, p)
    eng.reset()
    for k in range(1, cfg.T + 1):
        eng.rollout(1)
        goal = eng.get("goal_idx").reshape(cfg.N, M)
        force = eng.get("force").reshape(2, cfg.N, M)
        for n in range(cfg.N):
            out = sub(g, f"ped/{n}/out")
            assert np.array_equal(goal[n], out["goal"][k]), f"goal_idx scenario {n} tick {k}"
            close(force[:, n].T, out["force"][k], "social force")
    for n in range(cfg.N):
        check_against_golden(make_gpu, scene, p, sub(g, f"ped/{n}/out"), n, list(range(M)))


def test_road_network_golden():
    """Boundary forces among buildings + the ego_off_road terminal condition (general kernel)."""
    g = golden("road")
    cfg = golden_cases.ped_cfg()
    scene = pack_synthetic(cfg, road_network=golden_cases.road_network(golden_cases.ROAD_PED_GEOMETRY))
    p = _params(timestep=cfg.dt)
    M = cfg.M
    eng = make_gpu(scene, p)
    eng.reset()
    for k in range(1, cfg.T + 1):
        eng.rollout(1)
        goal = eng.get("goal_idx").reshape(cfg.N, M)
        force = eng.get("force").reshape(2, cfg.N, M)
        for n in range(cfg.N):
            out = sub(g, f"road_ped/{n}/out")
            assert np.array_equal(goal[n], out["goal"][k]), f"goal_idx scenario {n} tick {k}"
            close(force[:, n].T, out["force"][k], "social force with boundary forces")
    for n in range(cfg.N):
        check_against_golden(make_gpu, scene, p, sub(g, f"road_ped/{n}/out"), n, list(range(M)))
    cfg = golden_cases.veh_cfg()
    scene = pack_synthetic(cfg, road_network=golden_cases.road_network(golden_cases.ROAD_VEH_GEOMETRY))
    p = _params(timestep=cfg.dt, terminal=abi.TERM_MAX_LENGTH | abi.TERM_EGO_OFF_ROAD)
    for n in range(cfg.N):
        check_against_golden(make_gpu, scene, p, sub(g, f"road_veh/{n}/out"), n, list(range(cfg.M)),
                             actions=cfg.actions)
    # and on a large batch against the oracle: a crowd scene with a road network takes the general kernel
    big = synthetic.crowd_config(seed=21, N=2, M=300, T=6, side=6.0)
    scene = pack_synthetic(big, road_network=golden_cases.road_network(golden_cases.ROAD_PED_GEOMETRY))
    gpu, cpu = _run_both(scene, _params(timestep=big.dt))
    compare_engines(gpu, cpu, scene, "crowd with buildings", check_ped=True)


def test_pid_golden():
    from test_oracle_golden import pid_cases

    for name, spec, order, p, out in pid_cases():
        check_against_golden(make_gpu, pack_scenarios([spec]), p, out, 0, order)


def test_box_pairs_golden():
    """Exact closed-set box intersection incl. touching and identical boxes (unit vectors)."""
    import torch

    from scenario_gym_b200.abi import load_product

    lib = load_product()
    g = sub(golden("unit"), "unit")
    W, L, cx, cy = manifest()["unit"]["box"]
    n = len(g["pair_hit"])
    dev = torch.device("cuda:0")
    pa = torch.from_numpy(np.ascontiguousarray(g["pair_pose_a"][:, [0, 1, 3]])).to(dev)
    pb = torch.from_numpy(np.ascontiguousarray(g["pair_pose_b"][:, [0, 1, 3]])).to(dev)
    box = torch.tensor([[W, L, cx, cy]] * n, dtype=torch.float64, device=dev)
    out = torch.zeros(n, dtype=torch.uint8, device=dev)
    rc = lib["test_box_pairs"](pa.data_ptr(), box.data_ptr(), pb.data_ptr(), box.data_ptr(),
                               out.data_ptr(), n, 0, None)
    assert rc == 0
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), g["pair_hit"])


# ----------------------------------------------------------------------------- vs CPU oracle
def _run_both(scene, p, actions=None, n_calls=1, ticks=-1):
    gpu = make_gpu(scene, p)
    cpu = OracleEngine(scene, p, event_cap=1 << 16)
    for eng in (gpu, cpu):
        eng.reset()
        k0 = 0
        for _ in range(n_calls):
            if ticks < 0:
                eng.rollout(-1, actions=actions)
            else:
                eng.rollout(ticks, actions=None if actions is None else actions[k0: k0 + ticks])
                k0 += ticks
    return gpu, cpu


@pytest.mark.parametrize("seed", [0, 1])
def test_vehicles_vs_oracle(seed):
    """C3 shape (M = 64) at a size the oracle finishes in seconds; dense enough to collide."""
    cfg = synthetic.vehicles_config(seed=seed, N=96, M=64, T=48, half_extent=60.0)
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    gpu, cpu = _run_both(scene, p, actions=cfg.actions)
    assert cpu.get("n_pair_ticks").sum() > 100, "config must produce collisions"
    compare_engines(gpu, cpu, scene, f"vehicles seed {seed}")
    assert np.array_equal(gpu.get("coll_mask"), cpu.get("coll_mask"))
    # the same rollout split into several calls must give bit-identical results on the GPU
    gpu2, _ = _run_both(scene, p, actions=cfg.actions, n_calls=3, ticks=16)
    for k in ("pose", "vel", "dist", "speed", "t", "ego_avg_speed", "n_pair_ticks", "collided"):
        assert np.array_equal(gpu.get(k), gpu2.get(k)), f"split rollout differs in {k}"


@pytest.mark.parametrize("M", [1, 3, 20, 33, 100])
def test_ragged_slot_counts_vs_oracle(M):
    """Group sizes: sub-warp (several scenarios per warp), one warp, several warps with padding."""
    cfg = synthetic.vehicles_config(seed=7, N=37, M=M, T=24, half_extent=5.0 + 2.0 * M ** 0.5)
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    gpu, cpu = _run_both(scene, p, actions=cfg.actions)
    compare_engines(gpu, cpu, scene, f"M={M}")
    assert np.array_equal(gpu.get("coll_mask"), cpu.get("coll_mask"))


def test_highway_rss_vs_oracle():
    """C5 shape at reduced size: RSS + SafeDistance next to collisions."""
    cfg = synthetic.highway_config(seed=3, N=24, M=64, T=40, lanes=4)
    cfg.x0[:] = cfg.x0 * 0.5  # tighten headways so buffers are entered
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    p.features |= abi.FEAT_RSS
    gpu, cpu = _run_both(scene, p, actions=cfg.actions)
    assert cpu.get("rss_flags").any(), "config must violate RSS somewhere"
    compare_engines(gpu, cpu, scene, "highway", check_rss=True)


@pytest.mark.parametrize("M,calls,ticks", [(64, 1, -1), (64, 3, 14), (256, 1, -1), (256, 2, 20)])
def test_lean_rss_rollout_vs_oracle(M, calls, ticks):
    """
    Without a trace and a pair matrix the vehicle kernel runs its lean variant (the safe ratios,
    pure outputs, are evaluated once after the last tick of a call; M = 256 also takes the
    sorted sweep): whole and partial rollouts against the oracle.
    """
    cfg = synthetic.highway_config(seed=7, N=6, M=M, T=40, lanes=4)
    cfg.x0[:] = cfg.x0 * 0.5
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    p.features = abi.FEAT_COLLISIONS | abi.FEAT_EGO_METRICS | abi.FEAT_RSS
    gpu, cpu = _run_both(scene, p, actions=cfg.actions, n_calls=calls, ticks=ticks)
    assert cpu.get("rss_flags").any()
    compare_engines(gpu, cpu, scene, f"lean rss M={M}", check_rss=True)


def test_crowd_vs_oracle():
    """C4 shape at reduced size: social force with many neighbours per pedestrian."""
    cfg = synthetic.crowd_config(seed=5, N=6, M=96, T=30, side=9.0)
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    gpu, cpu = _run_both(scene, p)
    compare_engines(gpu, cpu, scene, "crowd", check_ped=True)


def test_big_group_vs_oracle():
    """M > 256: one scenario per CTA (M = 320)."""
    cfg = synthetic.crowd_config(seed=6, N=3, M=320, T=12, side=16.0)
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    gpu, cpu = _run_both(scene, p)
    compare_engines(gpu, cpu, scene, "crowd M=320", check_ped=True)


def _final_arrays(eng):
    keys = ("tick", "done", "present", "collided", "first_coll_tick", "first_coll_pair", "n_pair_ticks",
            "ego_hits", "t", "pose", "vel", "dist", "speed", "goal_idx", "force", "ego_avg_speed")
    return {k: eng.get(k).copy() for k in keys}


@pytest.mark.parametrize("case", ["sparse", "dense", "wrap", "large"])
def test_crowd_cell_grid(case):
    """
    Crowd scenarios with one CTA per scenario bin their entities into a shared-memory cell grid
    (sensor neighbours + collision broad phase).  The grid must change nothing: bit-identical to the
    exhaustive sweeps (FEAT_NO_GRID) and within tolerance of the CPU oracle.
      dense: more candidates than the per-pedestrian list holds; wrap: scene wider than the 64-cell
      torus; large: several vehicles too big for a cell walk through the crowd.
    """
    if case == "sparse":
        cfg = synthetic.crowd_config(seed=11, N=3, M=1024, T=10, side=40.0)
    elif case == "dense":
        cfg = synthetic.crowd_config(seed=12, N=2, M=288, T=8, side=5.0)
    elif case == "wrap":
        cfg = synthetic.crowd_config(seed=13, N=2, M=512, T=10, side=150.0)
        # clusters one torus period (64 cells of 1 m) apart alias into the same cells
        cfg.x0[:, 1::2] = cfg.x0[:, 1::2] % 8.0
        cfg.x0[:, 2::2] = cfg.x0[:, 2::2] % 8.0 + 64.0
        cfg.y0[:, 1:] = cfg.y0[:, 1:] % 12.0
    else:
        cfg = synthetic.crowd_config(seed=14, N=2, M=400, T=12, side=14.0)
        for k in range(1, 9):  # pedestrians with car-sized boxes standing in the crowd
            cfg.box[:, k] = synthetic.CAR1_BOX
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    gpu, cpu = _run_both(scene, p)
    compare_engines(gpu, cpu, scene, f"crowd grid/{case}", check_ped=True)
    assert int(cpu.get("n_pair_ticks").sum()) > 0, "case must exercise the broad phase"
    q = _params(timestep=cfg.dt)
    q.features |= abi.FEAT_NO_GRID
    ref = make_gpu(scene, q)
    ref.reset()
    ref.rollout(-1)
    a, b = _final_arrays(gpu), _final_arrays(ref)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], b[k]), \
            f"grid vs exhaustive sweep: {k} differs"


_RP_DISC = ("tick", "done", "present", "collided", "first_coll_tick", "first_coll_pair", "n_pair_ticks",
            "ego_hits", "t", "prev_t", "pose", "vel", "ego_avg_speed", "ego_max_speed", "coll_mask")


@pytest.mark.parametrize("terminal", ["max_length", "collision", "ego_collision"])
@pytest.mark.parametrize("calls", [(-1,), (50, 37, -1), (130, 8, 300, -1)])
@pytest.mark.parametrize("persist", [0, 1])
def test_replay_tick_parallel_equals_sequential(terminal, calls, persist):
    """
    Replay-only scenes are rolled out by the tick-parallel kernel (one thread per tick); it must give
    the sequential kernel's (FEAT_SEQUENTIAL) results bit for bit -- poses, velocities, tick times,
    collisions, events, ego metrics -- for whole and partial rollouts and for every terminal
    condition; only the accumulated distances are summed in a different (fixed) order.
    """
    specs = [s for _, s, _, _ in XOSC] + [s for _, s, _, _ in all_xosc_specs("xosc_norelabel")]
    scene = pack_scenarios(specs)
    term = {"max_length": abi.TERM_MAX_LENGTH, "collision": abi.TERM_MAX_LENGTH | abi.TERM_COLLISION,
            "ego_collision": abi.TERM_MAX_LENGTH | abi.TERM_EGO_COLLISION}[terminal]
    engines = []
    for seq in (False, True):
        p = _params(terminal=term, persist=persist)
        if seq:
            p.features |= abi.FEAT_SEQUENTIAL
        eng = make_gpu(scene, p)
        eng.reset()
        for k in calls:
            eng.rollout(k)
        engines.append(eng)
    par, seq = engines
    assert bool(seq.get("done").all())
    if terminal != "max_length":
        assert int((seq.get("first_coll_tick") >= 0).sum()) > 0, "case must end some rollouts early"
    for k in _RP_DISC:
        assert np.array_equal(par.get(k), seq.get(k), equal_nan=True), f"{k} differs"
    close(par.get("dist"), seq.get("dist"), "dist")
    close(par.get("ego_dist"), seq.get("ego_dist"), "ego_dist")
    a, b = par.events(), seq.events()
    assert len(a) == len(b)
    for f in ("scenario", "tick", "slot", "t"):
        assert np.array_equal(a[f], b[f]), f"event {f} differs"


@pytest.mark.parametrize("M", [17, 32])
def test_replay_wide_scenes_parallel_vs_sequential_and_oracle(M):
    """
    Replay-only scenes of up to 32 slots (the tick-parallel kernel's limit: 200 KB of staged boxes and poses per
    CTA at 32): crossing traffic with ragged knot counts, late arrivals and leavers, many overlapping boxes.
    Tick-parallel == sequential kernel bit for bit (distances: summation order), and both agree with the oracle.
    """
    from scenario_gym_b200.packing import ScenarioSpec, SlotSpec

    rng = np.random.default_rng(100 + M)
    specs = []
    for n in range(5):
        slots = []
        for s in range(M - (n % 3)):  # (ragged: the batch pads to M slots)
            K = int(rng.integers(1, 9))
            t = np.sort(rng.uniform(0.0 if s < 2 else -1.0, 9.0, K))
            t[0] = 0.0 if s == 0 else t[0]
            traj = np.zeros((K, 7))
            traj[:, 0] = t
            p0, v = rng.uniform(-12, 12, 2), rng.uniform(-4, 4, 2)
            traj[:, 1:3] = p0 + np.outer(t, v) + rng.normal(0, 0.3, (K, 2))
            traj[:, 3] = rng.normal(0, 0.05, K)
            traj[:, 4] = np.arctan2(v[1], v[0]) + rng.normal(0, 0.2, K)
            kind = abi.KIND_AGENT_REPLAY if s == 0 else abi.KIND_REPLAY
            slots.append(SlotSpec(kind=kind, traj=traj, box=(float(rng.uniform(0.6, 2.2)), float(rng.uniform(0.7, 4.5)),
                                                             float(rng.uniform(0, 1.4)), 0.0)))
        specs.append(ScenarioSpec(slots=slots))
    scene = pack_scenarios(specs)
    assert scene.M == M
    engines = []
    for seq in (False, True):
        p = _params()
        p.timestep = 1.0 / 20.0
        if seq:
            p.features |= abi.FEAT_SEQUENTIAL
        eng = make_gpu(scene, p)
        eng.reset()
        eng.rollout(-1)
        engines.append(eng)
    par, seq = engines
    assert int(par.get("n_pair_ticks").sum()) > 100, "the case must have collisions"
    for k in _RP_DISC:
        assert np.array_equal(par.get(k), seq.get(k), equal_nan=True), f"{k} differs"
    close(par.get("dist"), seq.get("dist"), "dist")
    a, b = par.events(), seq.events()
    assert a.tobytes() == b.tobytes()
    p = _params()
    p.timestep = 1.0 / 20.0
    cpu = OracleEngine(scene, p)
    cpu.reset()
    cpu.rollout(-1)
    for k in ("tick", "done", "present", "collided", "first_coll_tick", "first_coll_pair", "n_pair_ticks", "ego_hits"):
        assert np.array_equal(par.get(k), cpu.get(k)), f"oracle: {k} differs"
    close(par.get("pose"), cpu.get("pose"), "pose")
    close(par.get("ego_avg_speed"), cpu.get("ego_avg_speed"), "ego_avg_speed")


def test_replay_padding_invariance():
    """
    The same scenarios packed into 9, 16 and 32 slots (more trailing empty slots): the tick-parallel kernel stops
    its loops at the last non-empty slot, so every per-scenario result and every live slot's rows are unchanged.
    """
    specs = [s for _, s, _, _ in XOSC]
    outs = []
    for M in (None, 16, 32):
        scene = pack_scenarios(specs, n_slots=M)
        eng = make_gpu(scene, _params())
        eng.reset()
        eng.rollout(-1)
        outs.append((scene, eng))
    scene0, e0 = outs[0]
    M0 = scene0.M
    for scene, eng in outs[1:]:
        for k in ("tick", "done", "first_coll_tick", "first_coll_pair", "n_pair_ticks", "ego_hits", "t", "prev_t",
                  "ego_avg_speed", "ego_max_speed", "ego_dist"):
            assert np.array_equal(eng.get(k), e0.get(k), equal_nan=True), (scene.M, k)
        for k in ("present", "collided", "dist"):
            a = eng.get(k).reshape(scene.N, scene.M)[:, :M0]
            assert np.array_equal(a, e0.get(k).reshape(scene0.N, M0), equal_nan=True), (scene.M, k)
        for k in ("pose", "vel"):
            a = eng.get(k).reshape(6, scene.N, scene.M)[:, :, :M0]
            assert np.array_equal(a, e0.get(k).reshape(6, scene0.N, M0), equal_nan=True), (scene.M, k)
        assert eng.events().tobytes() == e0.events().tobytes()


def test_future_collision_detector_golden():
    """sg_future_collisions (one launch per batch) == the reference's FutureCollisionDetector flags."""
    hits = check_future_collisions(lambda scene, p: make_gpu(scene, p), _params())
    assert hits > 50


def test_future_collisions_vs_oracle():
    """Look-ahead collisions on dense synthetic traffic, every slot as the sensor entity in turn."""
    cfg = synthetic.highway_config(seed=21, N=12, M=32, T=8)
    cfg.x0[:] = cfg.x0 * 0.4
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    gpu, cpu = make_gpu(scene, p), OracleEngine(scene, p)
    rng = np.random.default_rng(3)
    seen = 0
    for s in range(0, cfg.M, 5):
        t = rng.uniform(-1.0, 3.0, cfg.N)
        slot = np.full(cfg.N, s, np.int32)
        for h, ns in ((5.0, 10), (0.7, 4), (2.0, 1)):
            a, b = gpu.future_collisions(t, h, ns, slot), cpu.future_collisions(t, h, ns, slot)
            assert np.array_equal(a, b), (s, h, ns)
            seen += int(b.sum())
    assert seen > 0


@pytest.mark.parametrize("M,half_extent", [(160, 45.0), (256, 25.0), (256, 9.0)])
def test_sorted_sweep_random_order_vs_oracle(M, half_extent):
    """
    Vehicle scenes with more than 128 slots keep their boxes sorted by x and sweep a fixed window.
    Random placement: the slots start in random x order (the first tick repairs the whole order),
    dense enough that runs of x-overlapping boxes exceed the window and, at 25 m, that the
    candidate queue overflows.
    """
    cfg = synthetic.vehicles_config(seed=31, N=3, M=M, T=20, half_extent=half_extent)
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    gpu, cpu = _run_both(scene, p, actions=cfg.actions)
    assert cpu.get("n_pair_ticks").sum() > 1000
    compare_engines(gpu, cpu, scene, f"sorted sweep M={M}")
    p2 = _params(timestep=cfg.dt)
    p2.features &= ~abi.FEAT_COLL_MATRIX  # the lean kernel variant
    gpu2, _ = make_gpu(scene, p2), None
    gpu2.reset()
    gpu2.rollout(-1, actions=cfg.actions)
    for k in ("collided", "first_coll_tick", "first_coll_pair", "n_pair_ticks", "ego_hits", "tick"):
        assert np.array_equal(gpu2.get(k), cpu.get(k)), f"lean variant: {k} differs"


def test_c2_replicas_identical():
    """C2: replicas of the test scenarios are bit-identical copies => identical results per file."""
    specs = [s for _, s, _, _ in XOSC]
    scene = pack_scenarios(specs)
    tiled = tile_scene(scene, 5)
    p = _params()
    p.features &= ~abi.FEAT_COLL_MATRIX
    gpu = make_gpu(tiled, p)
    gpu.reset()
    gpu.rollout(-1)
    cpu = OracleEngine(scene, p, event_cap=1 << 16)
    cpu.reset()
    cpu.rollout(-1)
    N = scene.N
    for k in ("tick", "first_coll_tick", "n_pair_ticks", "t"):
        a = gpu.get(k).reshape(5, N)
        assert all(np.array_equal(a[0], a[r]) for r in range(5)), k
        assert np.array_equal(a[0], cpu.get(k)), k
    for k in ("ego_avg_speed", "ego_max_speed", "ego_dist"):
        a = gpu.get(k).reshape(5, N)
        assert all(np.array_equal(a[0], a[r]) for r in range(5)), k
        close(a[0], cpu.get(k), k)
    assert len(gpu.events()) == 5 * len(cpu.events())
    # golden tick counts / end times of SURVEY.md section 8c
    for n, (name, _, _, out) in enumerate(XOSC):
        assert gpu.get("tick")[n] == int(out["n_ticks"])
        assert gpu.get("t")[n] == float(out["t_end"])


def test_properties_full_size_sample():
    """
    Size-independent properties on a larger batch (no oracle): permuting the scenarios of a
    batch permutes the results bit-identically; n_pair_ticks is consistent with collided flags.
    """
    cfg = synthetic.vehicles_config(seed=11, N=2048, M=64, T=32, half_extent=80.0)
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    p.features &= ~abi.FEAT_COLL_MATRIX
    gpu = make_gpu(scene, p)
    gpu.reset()
    gpu.rollout(-1, actions=cfg.actions)
    perm = np.random.default_rng(0).permutation(cfg.N)
    cfg2 = synthetic.vehicles_config(seed=11, N=2048, M=64, T=32, half_extent=80.0)
    for name in ("x0", "y0", "h0", "v0"):
        setattr(cfg2, name, getattr(cfg, name)[perm])
    cfg2.actions = cfg.actions.reshape(cfg.T, 2, cfg.N, cfg.M)[:, :, perm].reshape(cfg.T, 2, -1).copy()
    gpu2 = make_gpu(pack_synthetic(cfg2), p)
    gpu2.reset()
    gpu2.rollout(-1, actions=cfg2.actions)
    for k in ("ego_avg_speed", "ego_dist", "first_coll_tick", "n_pair_ticks", "tick"):
        assert np.array_equal(gpu.get(k)[perm], gpu2.get(k)), k
    pose = gpu.get("pose").reshape(6, cfg.N, cfg.M)
    assert np.array_equal(pose[:, perm], gpu2.get("pose").reshape(6, cfg.N, cfg.M))
    col = gpu.get("collided").reshape(cfg.N, cfg.M)
    npt = gpu.get("n_pair_ticks")
    assert np.array_equal(col.any(axis=1), npt > 0)
    assert np.array_equal(gpu.get("first_coll_tick") >= 0, npt > 0)


@pytest.mark.parametrize("M,N,rss", [(256, 5, True), (300, 3, True), (300, 2, False)])
def test_large_vehicle_groups_vs_oracle(M, N, rss):
    """Vehicle kernel variants: one 256-thread CTA per scenario, and the 1024-thread variant."""
    cfg = synthetic.highway_config(seed=9, N=N, M=M, T=16, lanes=4 if M % 4 == 0 else 3)
    cfg.x0[:] = cfg.x0 * 0.5
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    if rss:
        p.features |= abi.FEAT_RSS
    gpu, cpu = _run_both(scene, p, actions=cfg.actions)
    compare_engines(gpu, cpu, scene, f"highway M={M}", check_rss=rss)


def test_candidate_queue_overflow_vs_oracle():
    """So dense that the AABB survivors exceed the per-scenario queue (4 G entries): the kernel
    falls back to testing every survivor in place; results must not change."""
    cfg = synthetic.vehicles_config(seed=12, N=5, M=64, T=6, half_extent=3.0)
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    gpu, cpu = _run_both(scene, p, actions=cfg.actions)
    assert cpu.get("n_pair_ticks").min() > 4 * 64, "config must overflow the queue"
    compare_engines(gpu, cpu, scene, "queue overflow")
    assert np.array_equal(gpu.get("coll_mask"), cpu.get("coll_mask"))


def test_step_done_and_zero_ticks():
    """ScenarioGym.step() has no is_done guard (step_done=1); rollout() stops at is_done."""
    cfg = synthetic.vehicles_config(seed=2, N=3, M=8, T=10, half_extent=30.0)
    scene = pack_synthetic(cfg)
    p = _params(timestep=cfg.dt)
    table = np.concatenate([cfg.actions, cfg.actions], axis=0)
    for make in (make_gpu, lambda s, q, trace_cap=0: OracleEngine(s, q)):
        eng = make(scene, p)
        eng.reset()
        eng.rollout(0, actions=table)
        assert (eng.get("tick") == 0).all()
        eng.rollout(-1, actions=table)
        assert (eng.get("tick") == cfg.T).all() and eng.get("done").all()
        eng.rollout(2, actions=table[cfg.T:])  # done: skipped
        assert (eng.get("tick") == cfg.T).all()
        eng.rollout(2, actions=table[cfg.T:], step_done=True)  # stepped anyway
        assert (eng.get("tick") == cfg.T + 2).all()
