"""CPU test of the N>1 path: world_size-2 gloo processes shard a batch and gather the records."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from oracle.runner import OracleEngine
    from scenario_gym_b200 import abi, synthetic
    from scenario_gym_b200.distributed import gather_records, init_from_env, pack_records, shard_range
    from scenario_gym_b200.packing import slice_scene

    init_from_env("gloo")
    cfg = synthetic.vehicles_config(seed=3, N=n_total, M=8, T=12, half_extent=12.0)
    scene = synthetic.pack_synthetic(cfg)
    lo, hi = shard_range(n_total, rank, world)
    p = abi.default_params()
    p.timestep = cfg.dt
    eng = OracleEngine(slice_scene(scene, lo, hi), p)  # host stand-in for the per-GPU engine
    eng.reset()
    acts = cfg.actions.reshape(cfg.T, 2, n_total, cfg.M)[:, :, lo:hi].reshape(cfg.T, 2, -1)
    eng.rollout(-1, actions=np.ascontiguousarray(acts))
    fields = {k: torch.from_numpy(np.asarray(eng.get(k)).copy()) for k in
              ("ego_avg_speed", "ego_max_speed", "ego_dist", "first_coll_tick", "first_coll_pair",
               "n_pair_ticks", "rss_flags", "tick", "t")}
    allrec = gather_records(pack_records(fields), n_total)
    if rank == 0:
        q.put(allrec.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [11, 12])
def test_shard_and_gather_gloo(n_total):
    sys.path.insert(0, REPO)
    from oracle.runner import OracleEngine, build_oracle
    from scenario_gym_b200 import abi, synthetic
    from scenario_gym_b200.distributed import RECORD_FIELDS, shard_range

    build_oracle()
    world = 2  # n_total = 11: uneven shards, 6 + 5 (padded collective); 12: equal shards (gathered in place)
    if n_total == 11:
        assert [shard_range(n_total, r, world) for r in range(world)] == [(0, 6), (6, 11)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    cfg = synthetic.vehicles_config(seed=3, N=n_total, M=8, T=12, half_extent=12.0)
    p = abi.default_params()
    p.timestep = cfg.dt
    eng = OracleEngine(synthetic.pack_synthetic(cfg), p)
    eng.reset()
    eng.rollout(-1, actions=cfg.actions)
    assert got.shape == (n_total, len(RECORD_FIELDS))
    assert np.array_equal(got[:, 0], eng.get("ego_avg_speed"))
    assert np.array_equal(got[:, 3], eng.get("first_coll_tick"))
    assert np.array_equal(got[:, 6], eng.get("n_pair_ticks"))
    assert np.array_equal(got[:, 8], eng.get("tick"))
