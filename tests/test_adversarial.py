"""
Knife-edge geometry decided by rational arithmetic (tests/adversarial.py) against the CPU oracle and
the CUDA engine: box pairs (touching, collinear edges, +-1 ulp, zero-area / zero-length boxes,
identical boxes), the 64-gon of the sensor's radius query, polygon membership of the road-network
surfaces; plus the Trajectory truth tables of the reference evaluated on the device.
"""
import ctypes as C
from fractions import Fraction as F

import numpy as np
import pytest

from adversarial import box_pair_cases, ngon_cases

from helpers import golden
from oracle.runner import load_oracle, oracle_cdll


def test_rational_decider_sanity():
    pa, ba, pb, bb, want, labels = box_pair_cases()
    d = dict(zip(labels, want))
    assert d["centred: corner to corner"] and not d["centred: corner to corner+ulp"] and d["centred: corner to corner-ulp"]
    assert not d["centred: identical"] and d["centred: identical + ulp"]
    assert d["segments crossing"] and not d["segments parallel"] and not d["two identical points"]
    assert len(labels) >= 70 and 20 < want.sum() < len(want) - 20


def test_oracle_box_pairs_adversarial(oracle_lib):
    pa, ba, pb, bb, want, labels = box_pair_cases()
    out = np.zeros(len(want), np.uint8)
    oracle_lib["test_box_pairs"](pa.ctypes.data, ba.ctypes.data, pb.ctypes.data, bb.ctypes.data, out.ctypes.data,
                                 len(want), 0, None)
    bad = [l for l, w, o in zip(labels, want, out) if bool(w) != bool(o)]
    assert not bad, bad


def test_oracle_ngon_adversarial():
    dll = oracle_cdll()
    dll.sgo_in_buffer.restype = C.c_int
    dll.sgo_in_buffer.argtypes = [C.c_double] * 5
    rows = ngon_cases()
    bad = [l for (x, y, r, qx, qy, w, l) in rows if bool(dll.sgo_in_buffer(x, y, r, qx, qy)) != w]
    assert not bad, bad
    assert sum(1 for r in rows if r[5]) > 20 and sum(1 for r in rows if not r[5]) > 20


def _polygon_cases():
    """(edges [E, 4], px, py, expected side) -- an L-shaped ring with a square hole, exact expectations."""
    ext = [(0.0, 0.0), (6.0, 0.0), (6.0, 2.0), (2.0, 2.0), (2.0, 5.0), (0.0, 5.0)]
    hole = [(0.5, 0.5), (1.5, 0.5), (1.5, 1.5), (0.5, 1.5)]
    edges = []
    for ring in (ext, hole):
        for k in range(len(ring)):
            edges.append(ring[k] + ring[(k + 1) % len(ring)])
    up = lambda v: float(np.nextafter(v, np.inf))  # noqa: E731
    dn = lambda v: float(np.nextafter(v, -np.inf))  # noqa: E731
    pts = [((1.0, 3.0), 1), ((3.0, 1.0), 1), ((3.0, 3.0), -1), ((1.0, 1.0), -1), ((0.5, 1.0), 0), ((1.5, 1.5), 0),
           ((2.0, 2.0), 0), ((2.0, 3.0), 0), ((up(2.0), 3.0), -1), ((dn(2.0), 3.0), 1), ((6.0, 1.0), 0),
           ((up(6.0), 1.0), -1), ((0.0, 0.0), 0), ((3.0, 2.0), 0), ((3.0, dn(2.0)), 1), ((3.0, up(2.0)), -1),
           ((-1.0, 2.0), -1), ((1.0, 2.0), 1), ((4.0, 5.0), -1), ((1.0, 5.0), 0), ((dn(0.5), 1.0), 1), ((up(0.5), 1.0), -1)]
    return np.array(edges), pts


def test_oracle_polygon_side():
    dll = oracle_cdll()
    dll.sgo_polygon_side.restype = C.c_int
    dll.sgo_polygon_side.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double]
    edges, pts = _polygon_cases()
    from scenario_gym_b200.road_network import PolygonArea

    host = PolygonArea([(0.0, 0.0), (6.0, 0.0), (6.0, 2.0), (2.0, 2.0), (2.0, 5.0), (0.0, 5.0)],
                       [[(0.5, 0.5), (1.5, 0.5), (1.5, 1.5), (0.5, 1.5)]])
    for (px, py), want in pts:
        assert dll.sgo_polygon_side(edges.ctypes.data, len(edges), px, py) == want, (px, py)
        assert host.point_side(px, py) == want, (px, py)


# ------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_box_pairs_adversarial():
    import torch

    from scenario_gym_b200 import abi

    lib = abi.load_product()
    pa, ba, pb, bb, want, labels = box_pair_cases()
    dev = [torch.from_numpy(a).cuda() for a in (pa, ba, pb, bb)]
    out = torch.zeros(len(want), dtype=torch.uint8, device="cuda")
    rc = lib["test_box_pairs"](dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(),
                               out.data_ptr(), len(want), 0, None)
    assert rc == 0
    got = out.cpu().numpy()
    bad = [l for l, w, o in zip(labels, want, got) if bool(w) != bool(o)]
    assert not bad, bad


@pytest.mark.gpu
def test_gpu_rollout_collisions_adversarial():
    """The same pairs through the tick kernels' narrow phase (SAT filter + exact path): two static boxes per scenario."""
    from scenario_gym_b200 import abi
    from scenario_gym_b200.engine import Engine
    from scenario_gym_b200.packing import ScenarioSpec, SlotSpec, pack_scenarios

    pa, ba, pb, bb, want, labels = box_pair_cases()
    specs = []
    for k in range(len(want)):
        ta = np.array([[0.0, pa[k, 0], pa[k, 1], 0, 0, 0, 0]])
        tb = np.array([[0.0, pb[k, 0], pb[k, 1], 0, 0, 0, 0]])
        specs.append(ScenarioSpec(slots=[SlotSpec(kind=abi.KIND_AGENT_REPLAY, traj=ta, box=tuple(ba[k])),
                                         SlotSpec(kind=abi.KIND_REPLAY, traj=tb, box=tuple(bb[k]))], length=0.25))
    scene = pack_scenarios(specs)
    p = abi.default_params()
    p.timestep = 0.1
    for feats in (abi.FEAT_COLLISIONS, abi.FEAT_COLLISIONS | abi.FEAT_SEQUENTIAL):
        p.features = feats
        eng = Engine(scene, p, device=0)
        eng.reset()
        eng.rollout(2)
        hit = eng.get("first_coll_tick") >= 0
        bad = [l for l, w, o in zip(labels, want, hit) if bool(w) != bool(o)]
        assert not bad, (feats, bad)


@pytest.mark.gpu
def test_gpu_ngon_adversarial():
    """The sensor's 64-gon predicate through sg_entities_in_radius: one static entity per query point."""
    from scenario_gym_b200 import abi
    from scenario_gym_b200.engine import Engine
    from scenario_gym_b200.packing import ScenarioSpec, SlotSpec, pack_scenarios

    rows = ngon_cases()
    specs = [ScenarioSpec(slots=[SlotSpec(kind=abi.KIND_AGENT_REPLAY, traj=np.array([[0.0, qx, qy, 0, 0, 0, 0]]))])
             for (_, _, _, qx, qy, _, _) in rows]
    eng = Engine(pack_scenarios(specs), abi.default_params(), device=0)
    eng.reset()
    got = eng.entities_in_radius([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows])[:, 0]
    bad = [r[6] for r, g in zip(rows, got) if bool(g) != r[5]]
    assert not bad, bad


@pytest.mark.gpu
def test_gpu_trajectory_truth_tables():
    """position_at_t / velocity_at_t of the reference (tests/golden/unit.npz) on the device: bit-exact."""
    import torch

    from scenario_gym_b200 import abi

    lib = abi.load_product()
    g = {k[5:]: v for k, v in golden("unit").items() if k.startswith("unit/")}
    ts = torch.from_numpy(np.ascontiguousarray(g["traj_ts"])).cuda()
    n = ts.numel()
    for data_key, tag in (("traj_data", "traj"), ("traj1_data", "traj1")):
        rows = torch.from_numpy(np.ascontiguousarray(g[data_key])).cuda()
        for mode in (0, 1, 2):
            key = f"{tag}_pos_mode{mode}"
            if key not in g:
                continue
            pos = torch.zeros((n, 6), dtype=torch.float64, device="cuda")
            ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
            vel = torch.zeros((n, 6), dtype=torch.float64, device="cuda")
            rc = lib["test_trajectory"](rows.data_ptr(), rows.shape[0], ts.data_ptr(), n, mode, pos.data_ptr(),
                                        ok.data_ptr(), vel.data_ptr(), 0, None)
            assert rc == 0, lib["last_error"]()
            want = g[key]
            absent = np.isnan(want).all(axis=1)
            assert np.array_equal(ok.cpu().numpy().astype(bool), ~absent), (tag, mode)
            assert np.array_equal(pos.cpu().numpy()[~absent], want[~absent]), (tag, mode)
            if tag == "traj":
                assert np.array_equal(vel.cpu().numpy(), g["traj_vel"]), "velocity_at_t"
