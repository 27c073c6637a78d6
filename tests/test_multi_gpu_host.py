"""N > 1 host logic on CPU: two gloo ranks shard the swarm's environments, build their initial conditions
independently and all-gather their rollout statistics (the path's only collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_envs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from multidronesim_b200 import dist as mdist
    from multidronesim_b200 import scenarios
    r, lr, w = mdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = mdist.env_shard(total_envs, rank, world)
    init = scenarios.cbf_swarm_init(hi - lo, 8, seed=3, env_offset=lo)
    # stand-in for the device statistics vector of this shard (layout include/mds_b200.h MDS_STAT_*)
    stats = torch.tensor([8.0 * (hi - lo), float(init.sum()), float(np.abs(init).max()), float(init[..., 2].min()), hi - lo, 2.0 * (hi - lo), rank, 0.0],
                         dtype=torch.float64)
    allst = mdist.gather_stats(stats)
    red = mdist.reduce_stats(allst)
    np.save(os.path.join(out_dir, f"init_{rank}.npy"), init)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), red.numpy())
        np.save(os.path.join(out_dir, "all.npy"), allst.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_gather(tmp_path):
    from multidronesim_b200 import dist as mdist
    from multidronesim_b200 import scenarios
    world, total = 2, 1001  # uneven split on purpose
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    whole = scenarios.cbf_swarm_init(total, 8, seed=3, env_offset=0)
    parts = [np.load(tmp_path / f"init_{r}.npy") for r in range(world)]
    # the shards reproduce the single-process initial conditions env for env
    assert np.array_equal(np.concatenate(parts, axis=0), whole)
    bounds = [mdist.env_shard(total, r, world) for r in range(world)]
    assert bounds[0] == (0, 501) and bounds[1] == (501, 1001)
    red, allst = np.load(tmp_path / "reduced.npy"), np.load(tmp_path / "all.npy")
    assert allst.shape == (2, 8)
    assert red[0] == 8.0 * total and red[4] == total                      # sums
    assert np.isclose(red[1], whole.sum())
    assert red[2] == max(np.abs(p).max() for p in parts)                   # max over ranks
    assert red[3] == min(p[..., 2].min() for p in parts)                   # min over ranks
    assert red[6] == 1.0


def test_env_shard_covers_everything():
    from multidronesim_b200 import dist as mdist
    for total in (1, 7, 8, 125000, 1000003):
        for world in (1, 2, 4, 8):
            spans = [mdist.env_shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_counter_normal_is_shard_invariant_and_normal():
    from multidronesim_b200.scenarios import counter_normal
    a = counter_normal(3, np.arange(5000), 24)
    b = counter_normal(3, np.arange(1234, 5000), 24)
    assert np.array_equal(a[1234:], b)
    assert abs(a.mean()) < 0.01 and abs(a.std() - 1.0) < 0.01
    assert not np.array_equal(a, counter_normal(4, np.arange(5000), 24))
