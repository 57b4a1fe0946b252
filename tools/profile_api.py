#!/usr/bin/env python
"""cProfile of ScenarioGym.set_scenarios / rollout / get_metrics on 4096 replicas of the xosc test scenarios (GPU box)."""
import cProfile, os, pstats, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from helpers import golden, manifest, sub
from scenario_gym_b200 import (abi, BoundingBox, CatalogEntry, CollisionMetric, EgoAvgSpeed, Entity, Pedestrian, Scenario,
                               ScenarioGym, Trajectory, Vehicle)
g, man = golden("xosc"), manifest()["xosc"]
names = sorted(man)
cls = {abi.ETYPE_VEHICLE: (Vehicle, "Vehicle"), abi.ETYPE_PEDESTRIAN: (Pedestrian, "Pedestrian")}
def build(name):
    inp = sub(g, f"xosc/{name}/in")
    ents = []
    for i in range(int(inp["n_entities"])):
        C_, ctype = cls.get(int(inp["etype"][i]), (Entity, "MiscObject"))
        ce = CatalogEntry(None, "entry", None, ctype, BoundingBox(*[float(v) for v in inp["box"][i]]))
        ents.append(C_(ce, trajectory=Trajectory(inp[f"traj{i}"]), ref=man[name]["refs"][i]))
    return Scenario(ents, name=name)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t0 = time.perf_counter()
scenarios = [build(names[k % len(names)]) for k in range(n)]
print("build Scenario objects %.3f s" % (time.perf_counter() - t0))
gym = ScenarioGym(metrics=[CollisionMetric(), EgoAvgSpeed()], device=0)
for what, fn in (("set_scenarios", lambda: gym.set_scenarios(scenarios)), ("rollout", gym.rollout), ("rollout2", gym.rollout),
                 ("get_metrics", gym.get_metrics)):
    pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable(); fn(); pr.disable()
    print(f"==== {what}: {time.perf_counter() - t0:.4f} s")
    pstats.Stats(pr).sort_stats("cumulative").print_stats(12)

# ---- per-tick host mode: a custom Metric the engine knows nothing about
from scenario_gym_b200 import Metric
class MaxEntities(Metric):
    name = "max_entities"
    def _reset(self, state): self.value = len(state.poses)
    def _step(self, state): self.value = max(self.value, len(state.poses))
    def get_state(self): return self.value
gym2 = ScenarioGym(metrics=[EgoAvgSpeed(), MaxEntities()], device=0)
gym2.set_scenarios(scenarios[:64])
gym2.rollout()
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable(); gym2.rollout(); pr.disable()
print(f"==== host-metric rollout (64 scenarios): {time.perf_counter() - t0:.4f} s")
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
