"""N>1 path on CPU: two gloo ranks, envs sharded by global env index, no
collective in the step path, one all_gather of the rollout statistics.  The
per-rank stepping is done by the oracle here (no GPU in this container); the
sharding / offset / gather logic is the one bench.py and the multi-GPU launcher
use."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, steps, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from bioimitation_gym_b200 import registry, sharding
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local, offset = sharding.shard(n_total, rank, world)
    spec, cm, ref, task = registry.build_env_tables("MuscleWalkingImitation2D-v0", {})
    rt = orc.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
    env = orc.OracleVecEnv(cm.tables, task, rt, n_local, seed=5, env_offset=offset)
    obs = [env.reset()]
    acts = np.random.default_rng(0).uniform(0, 1, (steps, n_total, 14))
    ret = np.zeros(n_local)
    stats = torch.zeros(16, dtype=torch.float64)
    for k in range(steps):
        o, r, d, t, reasons = env.step(acts[k, offset:offset + n_local])
        obs.append(o)
        ret += r
        stats[0] += n_local
        stats[1] += float(d.sum())
        stats[2] += float(r.sum())
    gathered = sharding.all_gather_stats(stats)      # the only collective, outside the step path
    np.save(os.path.join(out_dir, "obs_%d.npy" % rank), np.stack(obs))
    if rank == 0:
        np.save(os.path.join(out_dir, "stats.npy"), gathered.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_batch(tmp_path, oracle_lib):
    import torch.multiprocessing as mp
    from bioimitation_gym_b200 import registry
    n_total, steps, world = 12, 8, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, steps, str(tmp_path)), nprocs=world, join=True)
    # single-process run of the whole batch
    spec, cm, ref, task = registry.build_env_tables("MuscleWalkingImitation2D-v0", {})
    rt = oracle_lib.RefTables(ref["q"], ref["u"], ref["body_pos"], ref["com_pos"])
    env = oracle_lib.OracleVecEnv(cm.tables, task, rt, n_total, seed=5)
    obs = [env.reset()]
    acts = np.random.default_rng(0).uniform(0, 1, (steps, n_total, 14))
    tot_r = 0.0
    n_done = 0
    for k in range(steps):
        o, r, d, t, _ = env.step(acts[k])
        obs.append(o)
        tot_r += r.sum()
        n_done += int(d.sum())
    full = np.stack(obs)
    parts = [np.load(os.path.join(str(tmp_path), "obs_%d.npy" % r)) for r in range(world)]
    np.testing.assert_array_equal(np.concatenate(parts, axis=1), full)   # bit-identical: RNG keyed by global env index
    stats = np.load(os.path.join(str(tmp_path), "stats.npy"))
    assert stats.shape == (world, 16)
    assert stats[:, 0].sum() == n_total * steps and stats[:, 1].sum() == n_done
    assert stats[:, 2].sum() == pytest.approx(tot_r, rel=1e-12)


def test_shard_arithmetic():
    from bioimitation_gym_b200 import sharding
    for n, w in ((4096, 1), (1 << 20, 8), (10, 4), (7, 8)):
        spans = [sharding.shard(n, r, w) for r in range(w)]
        assert sum(s[0] for s in spans) == n
        off = 0
        for cnt, o in spans:
            assert o == off
            off += cnt
