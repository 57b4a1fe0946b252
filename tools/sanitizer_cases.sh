cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from scenario_gym_b200 import abi, synthetic
from scenario_gym_b200.engine import Engine
from scenario_gym_b200.packing import pack_scenarios
from oracle import golden_cases
from helpers import all_xosc_specs
def run(scene, p, actions=None):
    e = Engine(scene, p, device=0, trace_cap=8); e.reset(); e.rollout(3, actions=actions); e.rollout(-1, actions=None if actions is None else actions[3:]); e.synchronize(); return e
p = abi.default_params(); p.features = abi.FEAT_COLLISIONS|abi.FEAT_EGO_METRICS|abi.FEAT_RSS|abi.FEAT_COLL_MATRIX
cfg = golden_cases.rss_cfg(); p.timestep = cfg.dt
run(synthetic.pack_synthetic(cfg), p, cfg.actions)
cfg = synthetic.vehicles_config(3, N=9, M=70, T=12, half_extent=20.0); p.timestep = cfg.dt
run(synthetic.pack_synthetic(cfg), p, cfg.actions)
cfg = golden_cases.ped_cfg(); p.timestep = cfg.dt; p.features &= ~abi.FEAT_RSS
run(synthetic.pack_synthetic(cfg), p)
cfg = synthetic.crowd_config(5, N=2, M=300, T=6, side=12.0); p.timestep = cfg.dt
run(synthetic.pack_synthetic(cfg), p)
specs = [s for _, s, _, _ in all_xosc_specs("xosc")][:6]
p = abi.default_params(); p.features |= abi.FEAT_COLL_MATRIX
e = Engine(pack_scenarios(specs), p, device=0); e.reset(); e.rollout(40); e.synchronize()
print("sanitizer cases done")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|done|Error" | head -12
done
