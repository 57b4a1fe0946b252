"""
TEST INFRASTRUCTURE ONLY -- the synthetic configurations behind tests/golden/*.npz,
shared by ``oracle/gen_golden.py`` (reference side) and the parity tests (engine side)
so both consume bit-identical inputs.
"""
import numpy as np

from scenario_gym_b200 import synthetic


def veh_cfg():
    """Dense random-action vehicles: 6 scenarios x 12 entities x 80 ticks in a 28 m square."""
    return synthetic.vehicles_config(seed=1, N=6, M=12, T=80, half_extent=14.0, name="veh")


def rss_cfg():
    """Highway-like traffic made eventful: an oncoming vehicle, a stationary one, tight headways."""
    cfg = synthetic.highway_config(seed=2, N=4, M=12, T=60, lanes=3, name="rss")
    cfg.h0[:, 5] += np.pi
    cfg.v0[:, 7] = 0.0
    cfg.x0[:] = cfg.x0 * 0.45
    cfg.actions[:, 0] *= 2.0
    return cfg


def ped_cfg():
    """Dense social-force crowd: 3 scenarios x (1 ego + 13 pedestrians) x 60 ticks, 5 m square."""
    return synthetic.crowd_config(seed=4, N=3, M=14, T=60, side=5.0)
