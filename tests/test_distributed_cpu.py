"""CPU tier, world_size = 2 over gloo: the N > 1 host logic (contiguous sharding by global env
index, optional statistics all-reduce).  Each rank steps its shard with the plain-C++ build of the
kernel source; gathered shards must equal the unsharded run bit for bit (RNG keyed by the global
index => env i is the same wherever it lives), and the reduced statistics must equal the global ones."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emul_harness as E
import helpers as H
from gym_pvder_b200 import _cabi
from gym_pvder_b200.sharding import STATS_FIELDS, reduce_stats, shard_bounds
from oracle import twin

TOTAL, STEPS, SEED = 37, 4, 21
KW = dict(model_type="model_1", events_spec=H.SAG_SPEC, seed=SEED, n_sim_time_steps_per_env_step=60, max_sim_time=3.0)


def _local_stats(env):
    ns = env.ns
    out = torch.zeros(16, dtype=torch.float64)
    out[0] = env.sd[_cabi.sd_index(ns)["ep_return"]].sum()
    out[1] = env.si[_cabi.SI_STEPS].sum()
    out[2] = (env.si[_cabi.SI_DONE] != 0).sum()
    out[3] = (env.si[_cabi.SI_STATUS] != 0).sum()
    for a in range(5):
        out[4 + a] = env.si[_cabi.SI_HIST + a].sum()
    out[9] = env.si[_cabi.SI_WINDUP].sum()
    out[10] = env.n
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(TOTAL, rank, world)
    env = E.EmulVecEnv(hi - lo, env_offset=lo, **KW)
    env.reset()
    for s in range(STEPS):
        a = twin.sample_actions_twin(SEED, s, TOTAL, 0)[lo:hi]
        env.step(a)
    stats = reduce_stats(_local_stats(env))
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, env.sd.copy(), env.si.copy()))
    if rank == 0:
        q.put((stats, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = E.EmulVecEnv(TOTAL, **KW)
    whole.reset()
    for s in range(STEPS):
        whole.step(twin.sample_actions_twin(SEED, s, TOTAL, 0))
    for lo, hi, sd, si in gathered:
        np.testing.assert_array_equal(sd, whole.sd[:, lo:hi])
        np.testing.assert_array_equal(si, whole.si[:, lo:hi])
    ref = _local_stats(whole).tolist()
    assert stats == dict(zip(STATS_FIELDS, ref[:len(STATS_FIELDS)])) or \
        all(abs(stats[k] - v) <= 1e-9 * max(1.0, abs(v)) for k, v in zip(STATS_FIELDS, ref))
    assert stats["n_envs"] == TOTAL and stats["steps_sum"] == 3 * TOTAL and stats["n_done"] == TOTAL
