"""CPU, world_size 2 over gloo: the N>1 host logic -- contiguous env-id shards, independent
per-shard stepping (no data-path collective), and the one collective of the path, the int64[8]
episode-statistics all-reduce.  The per-rank compute stand-in is the C oracle (tests may use it);
the sharding helpers and the all-reduce wrapper are the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, steps, seed, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle")]
    import c_oracle
    import py_oracle as po
    from gym_treasure_game_b200.vector_env import all_reduce_stats, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    b = c_oracle.CBatch(c_oracle.CLevel(po.default_level()), hi - lo, first_env_id=lo, seed=seed,
                        max_episode_steps=30, auto_reset=True)
    b.reset()
    rng = np.random.default_rng(123)                       # same global action stream on every rank
    rewards = []
    for _ in range(steps):
        a = rng.integers(0, 9, total, dtype=np.int32)
        _, r, _, _, _ = b.step(a[lo:hi])
        rewards.append(r.copy())
    stats = torch.from_numpy(b.stats().copy())
    all_reduce_stats(stats)                                # gloo here, NCCL on the GPUs
    np.save(os.path.join(out_dir, "rew%d.npy" % rank), np.stack(rewards))
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), stats.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_and_stats_allreduce(tmp_path):
    import c_oracle
    import py_oracle as po
    total, steps, seed, world = 1001, 60, 42, 2
    mp.start_processes(_worker, args=(world, _free_port(), total, steps, seed, str(tmp_path)), nprocs=world,
                       join=True, start_method="spawn")
    whole = c_oracle.CBatch(c_oracle.CLevel(po.default_level()), total, first_env_id=0, seed=seed,
                            max_episode_steps=30, auto_reset=True)
    whole.reset()
    rng = np.random.default_rng(123)
    ref = []
    for _ in range(steps):
        _, r, _, _, _ = whole.step(rng.integers(0, 9, total, dtype=np.int32))
        ref.append(r.copy())
    got = np.concatenate([np.load(tmp_path / ("rew%d.npy" % r)) for r in range(world)], axis=1)
    assert np.array_equal(got, np.stack(ref))                         # shards == whole population
    assert np.load(tmp_path / "stats.npy").tolist() == whole.stats().tolist()   # all-reduced sum == global stats
