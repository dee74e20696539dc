"""world_size-2 gloo tests of the N>1 plumbing (CPU): env sharding is independent of the number of ranks
(global-id keyed Philox) and the rollout-statistic reduction equals the single-process statistic."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import CC_TRACK, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from rui_b200 import abi
    from rui_b200.dist import allreduce_episode_stats, allreduce_moments, shard_range
    from rui_b200.env import packed_model

    off, n = shard_range(total, rank, world)
    pk = packed_model(True)
    cfg = abi.make_config(n, CC_TRACK, control_freq=500, torso_solref_randomization=True, initial_probe_pos_randomization=True, seed=3,
                          env_id_offset=off)
    # each rank simulates its own slice with the CPU oracle standing in for the GPU kernels (same Philox keying)
    obs = []
    for i in range(n):
        e = O.OracleEnv(pk, cfg, cfg.env_id_offset + i)
        obs.append(e.reset())
    obs = torch.tensor(np.array(obs))
    cnt = torch.tensor(float(n))
    mean = obs.mean(0)
    m2 = ((obs - mean) ** 2).sum(0)
    gc, gm, gm2 = allreduce_moments(cnt, mean, m2)
    rs, ls, ne = allreduce_episode_stats(torch.tensor(float(rank + 1)), torch.tensor(10.0 * (rank + 1)), torch.tensor(1.0))
    q.put((rank, off, n, obs.numpy(), float(gc), gm.numpy(), gm2.numpy(), float(rs), float(ls), float(ne)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_reset_and_stat_reduction_match_single_process():
    total, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference over the same global env ids
    from oracle import oracle as O
    from rui_b200 import abi
    from rui_b200.env import packed_model
    pk = packed_model(True)
    cfg = abi.make_config(total, CC_TRACK, control_freq=500, torso_solref_randomization=True, initial_probe_pos_randomization=True, seed=3)
    ref = np.array([O.OracleEnv(pk, cfg, i).reset() for i in range(total)])
    got = np.concatenate([r[3] for r in res])
    assert [(r[1], r[2]) for r in res] == [(0, 3), (3, 3)]
    np.testing.assert_array_equal(got, ref)  # sharding does not change a single bit
    for r in res:
        assert r[4] == total
        np.testing.assert_allclose(r[5], ref.mean(0), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(r[6], ((ref - ref.mean(0)) ** 2).sum(0), rtol=1e-9, atol=1e-9)
        assert (r[7], r[8], r[9]) == (3.0, 30.0, 2.0)
