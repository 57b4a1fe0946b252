"""Shared parity-test machinery: golden loading, scene building and comparison."""
from __future__ import annotations

import json
import os
from typing import Dict, List, Tuple

import numpy as np

from scenario_gym_b200 import abi
from scenario_gym_b200.packing import PackedScene, ScenarioSpec, SlotSpec, pack_scenarios

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_cache: Dict[str, Dict[str, np.ndarray]] = {}


def golden(name: str) -> Dict[str, np.ndarray]:
    if name not in _cache:
        with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
            _cache[name] = {k: z[k] for k in z.files}
    return _cache[name]


def manifest() -> dict:
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


def sub(d: Dict[str, np.ndarray], prefix: str) -> Dict[str, np.ndarray]:
    prefix = prefix.rstrip("/") + "/"
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def xosc_spec(inp: Dict[str, np.ndarray], agent_kind: int = abi.KIND_AGENT_REPLAY
              ) -> Tuple[ScenarioSpec, List[int]]:
    """ScenarioSpec from golden scenario inputs; returns slot -> entity index map."""
    n = int(inp["n_entities"])
    is_agent = inp["is_agent"].astype(bool)
    order = [i for i in range(n) if is_agent[i]] + [i for i in range(n) if not is_agent[i]]
    slots = []
    for i in order:
        slots.append(
            SlotSpec(
                kind=agent_kind if is_agent[i] else abi.KIND_REPLAY,
                traj=inp[f"traj{i}"],
                box=tuple(inp["box"][i]),
                etype=int(inp["etype"][i]),
            )
        )
    ego = int(inp["ego"])
    spec = ScenarioSpec(slots=slots, ego_slot=order.index(ego), first_slot=order.index(0))
    return spec, order


def all_xosc_specs(variant: str = "xosc"):
    g = golden("xosc")
    names = sorted({k.split("/")[1] for k in g if k.startswith(variant + "/")})
    out = []
    for name in names:
        spec, order = xosc_spec(sub(g, f"{variant}/{name}/in"))
        spec.name = name
        out.append((name, spec, order, sub(g, f"{variant}/{name}/out")))
    return out


def pairs_from_mask(mask: np.ndarray, M: int) -> List[Tuple[int, int]]:
    """(a, b), a < b, from one scenario's [M][W] uint32 pair matrix."""
    out = []
    for a in range(M):
        for b in range(a + 1, M):
            if (int(mask[a, b >> 5]) >> (b & 31)) & 1:
                assert (int(mask[b, a >> 5]) >> (a & 31)) & 1, "pair matrix not symmetric"
                out.append((a, b))
    return out


def check_against_golden(make_engine, scene: PackedScene, params, out: Dict[str, np.ndarray],
                         n: int, order: List[int], actions=None, tol: float = 1e-9,
                         check_pairs: bool = True) -> None:
    """
    Run scenario `n` of `scene` (a) fully fused with a trace and (b) tick by tick with the
    pair matrix, and compare with the reference's golden record `out`.
    Poses / continuous metrics: 1e-9 absolute-or-relative (north star); presence, tick
    count, collision pairs, events: exact.
    """
    T = int(out["n_ticks"])
    M, N = scene.M, scene.N
    NM = N * M
    sl = np.array([n * M + s for s in range(len(order))])
    inv = np.argsort(np.array(order))  # entity index -> slot

    def close(a, b, what):
        a, b = np.atleast_1d(np.asarray(a, np.float64)), np.atleast_1d(np.asarray(b, np.float64))
        assert a.shape == b.shape, (what, a.shape, b.shape)
        both_nan = np.isnan(a) & np.isnan(b)
        err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
        err[both_nan] = 0.0
        assert np.all(err <= tol), f"{what}: max err {np.nanmax(err):.3e}"

    # (a) fused rollout with trace
    eng = make_engine(scene, params, trace_cap=T + 2)
    eng.reset()
    eng.rollout(-1, actions=actions)
    tick = eng.get("tick")
    assert int(tick[n]) == T, f"tick count {int(tick[n])} != {T}"
    assert bool(eng.get("done")[n])
    tr_t = eng.get("trace_t")[: T + 1, n]
    assert np.array_equal(tr_t, out["t"]), "tick times must be bit-identical (repeated addition)"
    tr_present = eng.get("trace_present")[: T + 1][:, sl][:, inv]
    assert np.array_equal(tr_present, out["present"]), "presence mismatch"
    tr_pose = eng.get("trace_pose")[: T + 1][:, :, sl][:, :, inv]  # [T+1, 6, M]
    tr_pose = np.transpose(tr_pose, (0, 2, 1))
    masked = np.where(out["present"][..., None].astype(bool), tr_pose, np.nan)
    close(np.nansum(masked, axis=1), out["pose_sum"], "pose_sum")
    close(masked[out["keep"]], out["pose"], "pose")
    fin_vel = np.transpose(eng.get("vel")[:, sl][:, inv], (1, 0))
    fin_vel = np.where(out["present"][-1][:, None].astype(bool), fin_vel, np.nan)
    close(fin_vel, out["vel"][-1], "final velocity")
    close(eng.get("dist")[sl][inv], out["dist"][-1], "final distance")
    for key, field in (("ego_avg_speed", "ego_avg_speed"), ("ego_max_speed", "ego_max_speed"),
                       ("ego_distance_travelled", "ego_dist")):
        if key in out:
            close(eng.get(field)[n], out[key], key)
    # collisions: first tick / pair, events, pair-tick count
    gp = out["pairs"]
    gp_slots = sorted((int(t), *sorted((int(inv[a]), int(inv[b])))) for t, a, b in gp)
    if check_pairs:
        assert int(eng.get("n_pair_ticks")[n]) == len(gp_slots)
        if gp_slots:
            ft = gp_slots[0][0]
            assert int(eng.get("first_coll_tick")[n]) == ft
            assert tuple(eng.get("first_coll_pair")[n]) == gp_slots[0][1:]
        else:
            assert int(eng.get("first_coll_tick")[n]) == -1
        collided = np.zeros(len(order), bool)
        for _, a, b in gp_slots:
            collided[a] = collided[b] = True
        assert np.array_equal(eng.get("collided")[sl].astype(bool), collided)
        ev = eng.events()
        ev = ev[ev["scenario"] == n]
        got = sorted((int(e["tick"]), int(e["slot"])) for e in ev)
        want = sorted((int(t), int(inv[j])) for t, j in out["ego_events"])
        assert got == want, f"ego collision events {got} != {want}"
        ev_t = {(int(e["tick"]), int(e["slot"])): float(e["t"]) for e in ev}
        for (t, j), tt in zip(out["ego_events"], out["ego_event_t"]):
            assert ev_t[(int(t), int(inv[j]))] == float(tt)
    final_fused = {k: eng.get(k).copy() for k in ("pose", "vel", "dist", "t", "tick", "ego_avg_speed")}

    # (a') fused rollout without a trace: replay-only scenes take the tick-parallel kernel, whose
    # results must equal the sequential kernel's bit for bit, except the accumulated distances
    # (fixed summation tree over the ticks: tolerance)
    disc = ("tick", "done", "present", "collided", "first_coll_tick", "first_coll_pair", "n_pair_ticks",
            "ego_hits", "t", "prev_t", "pose", "vel", "ego_avg_speed", "ego_max_speed", "speed")
    seq = {k: eng.get(k).copy() for k in disc + ("dist", "ego_dist")}
    seq_ev = eng.events()
    par = make_engine(scene, params, trace_cap=0)
    par.reset()
    par.rollout(-1, actions=actions)
    for k in disc:
        assert np.array_equal(par.get(k), seq[k], equal_nan=True), f"trace-free fused rollout: {k} differs"
    close(par.get("dist"), seq["dist"], "trace-free fused rollout: dist")
    close(par.get("ego_dist"), seq["ego_dist"], "trace-free fused rollout: ego_dist")
    pe = par.events()
    key = lambda e: sorted(zip(e["scenario"].tolist(), e["tick"].tolist(), e["slot"].tolist(), e["t"].tolist()))
    assert key(pe) == key(seq_ev), "trace-free fused rollout: events differ"

    # (b) tick by tick with the pair matrix
    eng = make_engine(scene, params, trace_cap=0)
    eng.reset()
    got_pairs = []
    keep = {int(k): i for i, k in enumerate(out["keep"])}
    for k in range(1, T + 1):
        eng.rollout(1, actions=None if actions is None else actions[k - 1: k])
        if check_pairs:
            mask = eng.get("coll_mask")[n]
            got_pairs += [(k, a, b) for a, b in pairs_from_mask(mask, len(order))]
        if k in keep:
            v = np.transpose(eng.get("vel")[:, sl][:, inv], (1, 0))
            v = np.where(out["present"][k][:, None].astype(bool), v, np.nan)
            close(v, out["vel"][keep[k]], f"velocity@{k}")
            close(eng.get("dist")[sl][inv], out["dist"][keep[k]], f"distance@{k}")
    if check_pairs:
        assert got_pairs == gp_slots, "per-tick collision pairs differ from the reference"
    assert bool(eng.get("done")[n])
    for k, v in final_fused.items():
        a, b = eng.get(k), v
        if a.ndim and a.shape[-1] == NM:
            a, b = a[..., sl], b[..., sl]
        elif a.ndim and a.shape[0] == N:
            a, b = a[n], b[n]
        assert np.array_equal(a, b, equal_nan=True), f"fused vs stepwise {k} differ"


def check_future_collisions(make_engine, params) -> int:
    """
    ``future_collisions`` of an engine against the reference's FutureCollisionDetector flags
    (tests/golden/future.npz: every 20th tick of the 23 test scenarios, horizons 5.0 and 1.5).
    One batched call per (query index, horizon) over all scenarios.  Returns the number of hits.
    """
    from scenario_gym_b200.packing import pack_scenarios

    g = golden("future")
    specs = all_xosc_specs("xosc")
    scene = pack_scenarios([s for _, s, _, _ in specs])
    eng = make_engine(scene, params)
    eng.reset()
    names = [n for n, _, _, _ in specs]
    times = [g[f"future/{n}/t"] for n in names]
    hits = 0
    for h in (5.0, 1.5):
        flags = [g[f"future/{n}/flag_h{h}"].astype(bool) for n in names]
        for q in range(max(len(t) for t in times)):
            t = np.array([tt[min(q, len(tt) - 1)] for tt in times])
            want = np.array([f[min(q, len(f) - 1)] for f in flags])
            got = eng.future_collisions(t, horizon=h, n_samples=10)
            assert np.array_equal(got, want), f"future collisions differ at query {q}, horizon {h}"
            hits += int(want.sum())
    # default arguments: the current state.t and the ego
    assert np.array_equal(eng.future_collisions(),
                          np.array([g[f"future/{n}/flag_h5.0"][0] for n in names], bool))
    return hits
