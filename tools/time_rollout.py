#!/usr/bin/env python
"""Device time of reset + rollout for a workload, no checks (kernel experiments):  python tools/time_rollout.py c3 [N]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
from scenario_gym_b200 import abi, synthetic
from scenario_gym_b200.engine import Engine
w = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else {"c3": 12500, "c5": 10000, "c4": 1000, "c2": 4096}[w]
if w == "c2":
    import bench
    from types import SimpleNamespace
    scene_c2, per, tk = bench.c2_scene(N)
    cfg = SimpleNamespace(dt=1.0 / 30.0, T=1 << 20, M=1, c2_steps=int(per[np.arange(N) % len(per)].sum()))
elif w == "c3":
    cfg = synthetic.vehicles_config(seed=0, N=N, M=64, T=256, dt=0.1, materialise=False)
elif w == "c5":
    cfg = synthetic.highway_config(seed=0, N=N, M=256, T=256, materialise=False)
else:
    cfg = synthetic.crowd_config(seed=0, N=N, M=1024, T=128, dt=1.0 / 15.0)
scene = scene_c2 if w == "c2" else synthetic.pack_synthetic(cfg)
p = abi.default_params()
p.timestep = cfg.dt
p.features = abi.FEAT_COLLISIONS | abi.FEAT_EGO_METRICS | (abi.FEAT_RSS if w not in ("c4", "c2") else 0)
p.max_ticks = cfg.T
if os.environ.get("SG_FEATURES"):
    p.features = int(os.environ["SG_FEATURES"])  # experiments: 1 collisions, 2 ego metrics (abi.FEAT_*)
eng = Engine(scene, p, device=0)
act = getattr(cfg, "action_rng", None) if w not in ("c4", "c2") else None
best = 1e9
for it in range(6):
    eng.reset()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.rollout(-1, actions=act)
    b.record()
    torch.cuda.synchronize()
    if it >= 2:
        best = min(best, a.elapsed_time(b))
steps = cfg.c2_steps if w == "c2" else int(eng.get("tick").sum()) * cfg.M
print(f"{w} N={N}: rollout {best:.3f} ms, {steps / best / 1e6 * 1e3 / 1e9 * 1e3:.4g}e9 entity-steps/s" if False else f"{w} N={N}: rollout {best:.3f} ms  {steps / (best * 1e-3):.4g} entity-steps/s")
