"""The N > 1 path on CPU: two gloo ranks, each with its own shard of environments (test-only CPU build of the kernel
code), no per-step collective; the sharded job reproduces the single-process job environment by environment and the
episode statistics reduce to the same totals."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

EX = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'examples')
CFG = os.path.join(EX, 'ur_high_5', 'ur_high_5_randomised.yaml')
N_PER_RANK, STEPS = 3, 6


def _run(env, stats, n, offset):
    from diy_gym_b200.utils import walk_dict
    outs = []
    g = torch.Generator().manual_seed(7)
    all_actions = [torch.rand((2 * N_PER_RANK, 12), generator=g) * 0.02 - 0.01 for _ in range(STEPS)]
    for k in range(STEPS):
        a = all_actions[k][offset:offset + n]
        action = {'ur5_l': {'controller': {'linear': a[:, 0:3], 'rotation': a[:, 3:6]}}, 'ur5_r': {'controller': {'linear': a[:, 6:9], 'rotation': a[:, 9:12]}}}
        obs, rew, term, _ = env.step(action)
        done = term | (env.step_counter >= 3)           # end episodes every 3 steps so that there is something to reduce
        stats.update(walk_dict(rew, sum), done)
        if bool(done.any()):
            env.reset(done)
        outs.append(env.world.state.clone())
    return torch.stack(outs)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from diy_gym_b200 import DIYGym
    from diy_gym_b200.distributed import EpisodeStats, shard_offset
    from tests.emul.world import factory
    off = shard_offset(N_PER_RANK)
    env = DIYGym(CFG, num_envs=N_PER_RANK, seed=11, env_id_offset=off, world_factory=factory())
    stats = EpisodeStats(N_PER_RANK, 'cpu')
    states = _run(env, stats, N_PER_RANK, off)
    total = stats.reduce()
    dist.barrier()
    q.put((rank, off, states.numpy(), total))
    dist.destroy_process_group()


def test_two_gloo_ranks_match_single_process():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from diy_gym_b200 import DIYGym
    from diy_gym_b200.distributed import EpisodeStats
    from tests.emul.world import factory
    env = DIYGym(CFG, num_envs=2 * N_PER_RANK, seed=11, env_id_offset=0, world_factory=factory())
    stats = EpisodeStats(2 * N_PER_RANK, 'cpu')
    ref = _run(env, stats, 2 * N_PER_RANK, 0).numpy()
    total = stats.reduce()
    assert results[0][1] == 0 and results[1][1] == N_PER_RANK
    sharded = np.concatenate([results[0][2], results[1][2]], axis=1)
    assert np.array_equal(sharded, ref)                       # bit-identical, environment by environment
    assert results[0][3] == results[1][3]                     # both ranks hold the job-wide totals
    assert results[0][3]['episodes'] == total['episodes'] > 0
    assert abs(results[0][3]['mean_return'] - total['mean_return']) < 1e-9
