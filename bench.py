#!/usr/bin/env python
"""
bench.py -- entity-steps/s of the batched rollout + collision path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 engine
    python bench.py --impl reference --steps K --warmup W     # CPU arm (oracle port, all cores)

A "step" is one full rollout (reset + T ticks) of this rank's batch of synthetic scenarios.
Default workload = BASELINE.json configs[2] per-GPU shard (C3): 12 500 scenarios x 64
VehicleController entities x 256 ticks of dt = 0.1 with random accel/steer actions,
CollisionMetric + EgoAvgSpeed/EgoMaxSpeed/EgoDistanceTravelled + RSSDistances/RSS.
Weak scaling: every GPU gets its own 12 500 scenarios (seed = rank); no per-tick
communication; one NCCL all-gather of the per-scenario records at the end.

value  : device-timed (CUDA events, max over ranks), inputs resident in HBM.
e2e    : the same metric through sg_rollout_host with pinned HOST buffers: scene + action
         table H2D and result D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from scenario_gym_b200 import abi, synthetic  # noqa: E402
from scenario_gym_b200.packing import slice_scene  # noqa: E402

METRIC = "entity-steps/s, batched rollout+collision"
UNIT = "entity-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--scenarios-per-gpu", type=int, default=0)
    ap.add_argument("--ticks", type=int, default=256)
    ap.add_argument("--no-rss", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_spec(args):
    if args.workload == "c3":
        n = args.scenarios_per_gpu or 12500
        return dict(name="C3", N=n, M=64, T=args.ticks, dt=0.1)
    if args.workload == "c2":
        return dict(name="C2", N=args.scenarios_per_gpu or 4096, M=9, T=0, dt=1.0 / 30.0)
    if args.workload == "c4":
        return dict(name="C4", N=args.scenarios_per_gpu or 1000, M=1024, T=min(args.ticks, 128), dt=1.0 / 15.0)
    n = args.scenarios_per_gpu or 1250
    return dict(name="C5", N=n, M=256, T=args.ticks, dt=0.1)


def features(args) -> int:
    f = abi.FEAT_COLLISIONS | abi.FEAT_EGO_METRICS
    if not args.no_rss and args.workload in ("c3", "c5"):
        f |= abi.FEAT_RSS
    return f


def algorithmic_bytes_per_entity_step(args) -> int:
    """SURVEY.md section 8d B_tick (per-tick streaming design, fp64 SoA)."""
    if args.workload == "c2":
        return 290
    if args.workload == "c4":
        return 305
    return 225 if args.no_rss else 259


def c2_scene(n_scen: int):
    """C2: replicas of the reference's 23 test scenarios (inputs committed under tests/golden)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from helpers import all_xosc_specs

    from scenario_gym_b200.packing import pack_scenarios, tile_scene

    specs = [s for _, s, _, _ in all_xosc_specs("xosc")]
    base = pack_scenarios(specs)
    reps = -(-n_scen // base.N)
    return slice_scene(tile_scene(base, reps), 0, n_scen)


def make_config(args, seed: int, n_scen: int, actions_out=None):
    w = workload_spec(args)
    if args.workload == "c4":
        return synthetic.crowd_config(seed=seed, N=n_scen, M=w["M"], T=w["T"], dt=w["dt"])
    if args.workload == "c3":
        return synthetic.vehicles_config(seed=seed, N=n_scen, M=w["M"], T=w["T"], dt=w["dt"],
                                         actions_out=actions_out)
    cfg = synthetic.highway_config(seed=seed, N=n_scen, M=w["M"], T=w["T"], dt=w["dt"])
    if actions_out is not None:
        actions_out[:] = cfg.actions
        cfg.actions = actions_out
    return cfg


def config_json(args, extra=None):
    w = workload_spec(args)
    mets = ["CollisionMetric", "EgoAvgSpeed", "EgoMaxSpeed", "EgoDistanceTravelled"]
    if features(args) & abi.FEAT_RSS:
        mets += ["RSSDistances", "RSS"]
    desc = {
        "c2": f"C2: {w['N']} replicas of the reference's 23 test scenarios (1-9 entities, 324-722 ticks at "
              "30 Hz), trajectory replay + CollisionMetric/ego metrics (BASELINE.json configs[1])",
        "c3": f"C3: {w['N']} scenarios/GPU x {w['M']} VehicleController entities x {w['T']} ticks, random "
              "accel/steer actions (BASELINE.json configs[2] per-GPU shard)",
        "c4": f"C4: {w['N']} scenarios x {w['M']} social-force pedestrians x {w['T']} ticks "
              "(BASELINE.json configs[3])",
        "c5": f"C5: {w['N']} scenarios/GPU x {w['M']} highway vehicles x {w['T']} ticks, RSS + SafeDistance "
              "(BASELINE.json configs[4] per-GPU shard)",
    }[args.workload]
    l2 = ("inputs larger than L2: the action table read by every step is "
          f"{w['N'] * w['M'] * w['T'] * 16 / 1e9:.2f} GB") if args.workload in ("c3", "c5") else \
        "L2 flushed between steps: the reset kernel rewrites every state plane before each rollout"
    c = {"workload": desc, "scenarios_per_gpu": w["N"], "entities": w["M"], "ticks": w["T"],
         "timestep": w["dt"], "metrics": mets, "l2_policy": l2}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed region.

    NVML is polled from a thread every few ms (the timed region of a short run is < 100 ms, too
    short for `nvidia-smi -lms`); `nvidia-smi` is the fall-back when NVML cannot be loaded.
    """

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, uuid: str = None, period_s: float = 0.004):
        self.index, self.uuid, self.period = index, uuid, period_s
        self.proc = None
        self.nvml = None
        self.samples = []  # (wall time, sm MHz, max MHz, watts, set of reasons)
        self._stop = threading.Event()
        self.thread = None

    # -- NVML ---------------------------------------------------------------------------
    def _nvml_open(self):
        import pynvml as nv

        nv.nvmlInit()
        h = None
        if self.uuid:
            for u in (self.uuid, "GPU-" + self.uuid):
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(u.encode() if isinstance(u, str) else u)
                    break
                except Exception:
                    h = None
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        self.nvml, self.handle = nv, h

    def _nvml_loop(self):
        nv, h = self.nvml, self.handle
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                try:
                    watts = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                except Exception:
                    watts = float("nan")
                self.samples.append((time.time(), mhz, self.max_mhz, watts,
                                     {nm for b, nm in bits if r & b}))
            except Exception:
                pass
            self._stop.wait(self.period)

    # -- nvidia-smi fall-back -------------------------------------------------------------
    def _smi_loop(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) < 7:
                continue
            try:
                self.samples.append((time.time(), float(parts[0]), float(parts[1]), float(parts[2]),
                                     {nm for nm, v in zip(self.NAMES, parts[3:7])
                                      if v.lower().startswith("active")}))
            except ValueError:
                continue

    def start(self):
        try:
            self._nvml_open()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def window(self, t0: float, t1: float) -> None:
        """Only samples taken inside [t0, t1] (the timed region) are reported."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML / nvidia-smi"]}
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is not None:
            self.thread.join(timeout=2)
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        reasons = set()
        for s in inside:
            reasons |= s[4]
        power = [s[3] for s in inside if s[3] == s[3]]
        return {
            "sm_mhz": float(np.median([s[1] for s in inside])) if inside else None,
            "sm_max_mhz": float(max(s[2] for s in inside)) if inside else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(inside),
            "source": "nvml" if self.nvml is not None else "nvidia-smi",
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------ CPU arms
_W = {}


def _worker_init(args_dict, per_worker):
    import argparse as _ap

    from oracle.runner import OracleEngine

    args = _ap.Namespace(**args_dict)
    ident = os.getpid()
    scene, p, actions, steps = cpu_sample(args, seed=1000 + ident % 1000, n_scen=per_worker)
    _W["eng"] = OracleEngine(scene, p, event_cap=1 << 16)
    _W["actions"], _W["steps"], _W["M"] = actions, steps, scene.M


def _worker_step(_):
    eng = _W["eng"]
    t0 = time.perf_counter()
    eng.reset()
    eng.rollout(-1, actions=_W["actions"])
    steps = _W["steps"] if _W["steps"] is not None else int(eng.get("tick").sum()) * _W["M"]
    return steps, time.perf_counter() - t0


def cpu_sample(args, seed: int, n_scen: int):
    """(scene, params, actions, entity-steps per rollout or None) of a bounded sample of the workload."""
    if args.workload == "c2":
        scene = c2_scene(n_scen)
        sys.path.insert(0, os.path.join(REPO, "tests"))
        from helpers import all_xosc_specs

        per = np.array([int(o["present"][1:].sum()) for _, _, _, o in all_xosc_specs("xosc")], np.int64)
        steps = int(per[np.arange(n_scen) % len(per)].sum())
        dt, actions = workload_spec(args)["dt"], None
    else:
        cfg = make_config(args, seed=seed, n_scen=n_scen)
        scene = synthetic.pack_synthetic(cfg)
        steps, dt, actions = None, cfg.dt, cfg.actions
    p = abi.default_params()
    p.timestep = dt
    p.features = features(args)
    return scene, p, actions, steps


def cpu_oracle_single(args, budget_s: float = 12.0):
    """The oracle port on ONE core over a bounded sample of the same workload."""
    from oracle.runner import OracleEngine

    w = workload_spec(args)
    n = 8
    while True:
        scene, p, actions, steps = cpu_sample(args, seed=0, n_scen=n)
        eng = OracleEngine(scene, p, event_cap=1 << 16)
        t0 = time.perf_counter()
        eng.reset()
        eng.rollout(-1, actions=actions)
        dt = time.perf_counter() - t0
        if steps is None:
            steps = int(eng.get("tick").sum()) * scene.M
        if dt >= budget_s / 4 or n >= 2048:
            break
        n = min(2048, max(n * 2, int(n * budget_s / max(dt, 1e-3) / 2)))
    return {
        "value": steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
        "sample": f"{n} scenarios x {w['M']} entities x {w['T'] or 'all'} ticks of the same workload "
                  f"({steps} entity-steps in {dt:.2f} s), oracle/sg_oracle.c single thread",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    from oracle.runner import build_oracle

    build_oracle()
    cores = os.cpu_count() or 1
    w = workload_spec(args)
    per_worker = max(1, int(round(4e5 / (w["M"] * max(w["T"], 200)))))  # ~0.2-0.4 s of work per step per core
    args_dict = vars(args)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init, initargs=(args_dict, per_worker)) as pool:
        for _ in range(args.warmup):
            pool.map(_worker_step, range(cores), chunksize=1)
        t0 = time.perf_counter()
        total = 0
        for _ in range(args.steps):
            res = pool.map(_worker_step, range(cores), chunksize=1)
            total += sum(r[0] for r in res)
        dt = time.perf_counter() - t0
    value = total / dt
    sample = (f"{cores} processes x {per_worker} scenarios x {w['M']} entities x {w['T']} ticks per step; "
              "plain-C port of the reference path (oracle/sg_oracle.c); the Python reference itself "
              "cannot travel to the GPU box (measured in the authoring container: ~1.5e4 entity-steps/s/core)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_json(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from scenario_gym_b200.distributed import gather_records, init_from_env, pack_records
    from scenario_gym_b200.engine import Engine

    rank, world, local = init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    w = workload_spec(args)
    N, M, T = w["N"], w["M"], w["T"]
    NM = N * M

    # inputs: generated straight into pinned host memory (the e2e path copies from there)
    act_host = act_dev = None
    expected_ticks = None
    if args.workload == "c2":
        scene = c2_scene(N)
        M = scene.M
        NM = N * M
        sys.path.insert(0, os.path.join(REPO, "tests"))
        from helpers import all_xosc_specs

        outs = [o for _, _, _, o in all_xosc_specs("xosc")]
        per = np.array([int(o["present"][1:].sum()) for o in outs], np.int64)
        tk = np.array([int(o["n_ticks"]) for o in outs], np.int64)
        idx = np.arange(N) % len(outs)
        steps_expected = int(per[idx].sum())
        expected_ticks = tk[idx]
        dt = w["dt"]
    else:
        if args.workload in ("c3", "c5"):
            act_host = torch.empty((T, 2, NM), dtype=torch.float64, pin_memory=True)
        cfg = make_config(args, seed=rank, n_scen=N,
                          actions_out=None if act_host is None else act_host.numpy())
        scene = synthetic.pack_synthetic(cfg)
        steps_expected = N * M * T
        expected_ticks = np.full(N, T)
        dt = cfg.dt
    l2_note = None
    if args.workload == "c2":
        in_bytes = sum(a.nbytes for a in scene.arrays().values())
        l2_note = (f"inputs larger than L2: the trajectory / union-knot tables read by every step are "
                   f"{in_bytes / 1e6:.0f} MB")
    p = abi.default_params()
    p.timestep = dt
    p.features = features(args)
    eng = Engine(scene, p, device=dev, event_cap=1 << 22)
    if act_host is not None:
        act_dev = eng.set_actions(act_host)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step():
        eng.reset()
        eng.rollout(-1, actions=act_dev)

    # ---- device-resident timing -----------------------------------------------------
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
    barrier()
    t_wall0 = time.time()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for k in range(args.steps):
        ev[k][0].record(stream)
        eng.reset()
        ev[k][1].record(stream)
        eng.rollout(-1, actions=act_dev)
        ev[k][2].record(stream)
    stop.record(stream)
    barrier()
    sampler.window(t_wall0, time.time())
    elapsed_ms = start.elapsed_time(stop)
    kern_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    reset_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    clocks = sampler.stop() if rank == 0 else None
    ticks = eng.get("tick")
    steps_per_rollout = steps_expected  # entity-steps = sum over ticks of entities present
    assert np.array_equal(ticks, expected_ticks), "every scenario must run its full number of ticks"

    tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    total_steps = torch.tensor([float(steps_per_rollout)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_steps, op=dist.ReduceOp.SUM)
    elapsed_ms = float(tmax.item())
    value = float(total_steps.item()) * args.steps / (elapsed_ms / 1e3)

    # ---- final metric gather (the only collective on the path) -------------------------
    barrier()
    g0 = time.perf_counter()
    fields = {k: eng.tensor(k) for k in ("ego_avg_speed", "ego_max_speed", "ego_dist", "first_coll_tick",
                                         "first_coll_pair", "n_pair_ticks", "rss_flags", "tick", "t")}
    records = gather_records(pack_records(fields), N * world)
    torch.cuda.synchronize(dev)
    gather_ms = (time.perf_counter() - g0) * 1e3
    assert records.shape[0] == N * world

    # ---- end to end through the host-buffer C-ABI call ---------------------------------
    e2e = None
    if not args.no_e2e:
        lib = eng.lib
        host_keep = {k: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                     for k, a in scene.arrays().items()}
        hs = abi.SgScene()
        for f, _ in abi.SgScene._fields_:
            setattr(hs, f, getattr(eng._sc, f))
        for k, t in host_keep.items():
            setattr(hs, k, t.data_ptr() if t.numel() else None)
        hin, din = abi.SgInputs(), abi.SgInputs()
        if act_host is not None:
            hin.actions, hin.n_action_ticks = act_host.data_ptr(), T
            din.actions, din.n_action_ticks = act_dev.data_ptr(), T
        res_keep = {
            "ego_avg_speed": torch.empty(N, dtype=torch.float64, pin_memory=True),
            "ego_max_speed": torch.empty(N, dtype=torch.float64, pin_memory=True),
            "ego_dist": torch.empty(N, dtype=torch.float64, pin_memory=True),
            "first_coll_tick": torch.empty(N, dtype=torch.int32, pin_memory=True),
            "first_coll_pair": torch.empty((N, 2), dtype=torch.int32, pin_memory=True),
            "n_pair_ticks": torch.empty(N, dtype=torch.int64, pin_memory=True),
            "rss_flags": torch.empty(N, dtype=torch.uint8, pin_memory=True),
            "tick": torch.empty(N, dtype=torch.int32, pin_memory=True),
            "t": torch.empty(N, dtype=torch.float64, pin_memory=True),
            "event_count": torch.empty(1, dtype=torch.int32, pin_memory=True),
        }
        res = abi.SgHostResults()
        for k, t in res_keep.items():
            setattr(res, k, t.data_ptr())
        h2d = int(lib["host_h2d_bytes"](C.byref(hs), C.byref(hin), 1))
        d2h = int(lib["host_d2h_bytes"](C.byref(hs)))

        def e2e_step():
            rc = lib["rollout_host"](C.byref(hs), C.byref(eng._sc), C.byref(p), C.byref(eng._st),
                                     C.byref(hin), C.byref(din), C.byref(res), 1, eng.dev_index,
                                     stream.cuda_stream)
            if rc:
                raise RuntimeError(lib["last_error"]().decode())

        ref_avg = eng.get("ego_avg_speed").copy()
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
            stream.synchronize()  # the caller reads the results after every rollout
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        assert np.array_equal(res_keep["ego_avg_speed"].numpy(), ref_avg), "e2e path result mismatch"
        assert np.array_equal(res_keep["tick"].numpy(), expected_ticks)
        tw = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e = {"value": float(total_steps.item()) * args.steps / float(tw.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * float(tw.item()) / args.steps}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (sg_rollout_kernel) -----------------------------
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bpe = algorithmic_bytes_per_entity_step(args)
    achieved = steps_per_rollout * bpe / (kern_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        key = f"{args.workload}{'' if not args.no_rss else '_norss'}_{N}x{M}x{T}"
        traffic = tj.get(key)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "kernel": {"c3": "sg_vehicle_kernel<RSS=%d>" % (0 if args.no_rss else 1),
                                       "c5": "sg_vehicle_kernel<RSS=%d>" % (0 if args.no_rss else 1),
                                       "c2": "sg_replay_kernel (tick-parallel)",
                                       "c4": "sg_rollout_kernel<PED=1> (cell grid)"}[args.workload],
        "kernel_ms": kern_ms, "reset_kernel_ms": reset_ms,
        "algorithmic_bytes_per_entity_step": bpe, "entity_steps_per_launch": steps_per_rollout,
        "peak_source": peak_src,
        "note": "achieved = SURVEY 8d per-tick-streaming bytes (B_tick) x entity-steps / kernel time, "
                "of measured HBM copy bandwidth; the fused kernels keep State rows on chip across "
                "ticks, so their real DRAM traffic (traffic, from ncu) is far below B_tick and they "
                "are issue/latency bound, not HBM bound (profiles/)",
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle.runner import build_oracle

        build_oracle()
        cpu = cpu_oracle_single(args)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_json(args, dict({"parallelism": f"scenario-sharded x{world}, no per-tick communication"},
                                         **({"l2_policy": l2_note} if l2_note else {}))),
        "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps, "roofline": roofline,
        "cpu_baseline": cpu, "gather_ms": gather_ms,
        "collisions": {"pair_ticks": int(eng.get("n_pair_ticks").sum()),
                       "scenarios_with_collision": int((eng.get("first_coll_tick") >= 0).sum()),
                       "ego_events": int(eng.tensor("event_count").item())},
    }
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
