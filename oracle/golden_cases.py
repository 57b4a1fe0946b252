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


# ---------------------------------------------------------------------------- road networks
# plain coordinate lists: the reference side wraps them in shapely polygons (oracle/gen_golden.py),
# the engine side in scenario_gym_b200.road_network objects
ROAD_PED_GEOMETRY = {
    # pedestrians walk in the 5 m square [0, 5]^2 (ped_cfg); two buildings stand in it, one pavement covers it
    "buildings": [
        [(1.8, 1.9), (2.9, 1.9), (2.9, 2.7), (1.8, 2.7)],
        [(3.6, 0.4), (4.8, 0.4), (4.8, 1.0), (4.2, 1.0), (4.2, 1.6), (3.6, 1.6)],  # L-shaped
    ],
    "pavements": [
        {"exterior": [(-2.0, -2.0), (7.0, -2.0), (7.0, 7.0), (-2.0, 7.0)],
         "interiors": [[(0.2, 3.9), (0.9, 3.9), (0.9, 4.6), (0.2, 4.6)]]},
    ],
    "roads": [[(-2.0, -14.0), (7.0, -14.0), (7.0, -6.0), (-2.0, -6.0)]],
}
ROAD_VEH_GEOMETRY = {
    # vehicles start in the 28 m square [-14, 14]^2 (veh_cfg); the driveable surface is an L-shaped
    # road plus a crossing road sharing an edge with it: egos leave it at different ticks
    "roads": [
        [(-16.0, -16.0), (16.0, -16.0), (16.0, 2.0), (3.0, 2.0), (3.0, 16.0), (-16.0, 16.0)],
        [(16.0, -6.0), (40.0, -6.0), (40.0, 2.0), (16.0, 2.0)],
    ],
    "buildings": [],
    "pavements": [],
}


def road_network(geometry):
    """The geometry as a scenario_gym_b200 RoadNetwork (engine side)."""
    from scenario_gym_b200.road_network import Building, Pavement, PolygonArea, Road, RoadNetwork

    def poly(b):
        return PolygonArea.from_json(b) if isinstance(b, dict) else PolygonArea(b)

    return RoadNetwork(
        roads=[Road(f"road_{k}", poly(b)) for k, b in enumerate(geometry["roads"])],
        intersections=[],
        pavements=[Pavement(f"pavement_{k}", poly(b)) for k, b in enumerate(geometry["pavements"])],
        buildings=[Building(f"building_{k}", poly(b)) for k, b in enumerate(geometry["buildings"])],
    )
