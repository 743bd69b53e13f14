"""Multi-GPU plumbing: environments shard across ranks with NO per-step collective.

Rank r of W owns the global environment ids [r * n, (r + 1) * n): pass `env_id_offset=shard_offset(n)` to DIYGym so
that every per-environment random stream is keyed by the global id and results do not depend on the GPU count.
The only collective is the optional episode-statistics reduction below (NCCL over NVLink on GPUs, gloo in CPU tests);
it moves a handful of floats and is off the step path.
"""
import os

import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def shard_offset(envs_per_rank):
    """Global id of this rank's environment 0."""
    return rank_world()[0] * int(envs_per_rank)


class EpisodeStats:
    """Running per-rank sums of finished episodes; `reduce()` returns the job-wide totals."""
    def __init__(self, num_envs, device):
        self.ret = torch.zeros(num_envs, device=device)
        self.length = torch.zeros(num_envs, device=device)
        self.sums = torch.zeros(4, dtype=torch.float64, device=device)   # episodes, sum return, sum return^2, sum length

    def update(self, reward, done):
        """reward: [N] tensor (already summed over add-ons), done: [N] bool tensor."""
        self.ret += reward
        self.length += 1
        # no host synchronisation on the step path: finished episodes are folded in by masked arithmetic
        d = done.to(self.ret.dtype)
        r, l = (self.ret * d).double(), (self.length * d).double()
        self.sums += torch.stack([d.double().sum(), r.sum(), (r * r).sum(), l.sum()])
        self.ret *= 1 - d
        self.length *= 1 - d

    def reduce(self):
        total = self.sums.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM)
        n = max(float(total[0]), 1.0)
        mean = float(total[1]) / n
        return dict(episodes=int(total[0]), mean_return=mean, var_return=max(float(total[2]) / n - mean * mean, 0.0), mean_length=float(total[3]) / n)
