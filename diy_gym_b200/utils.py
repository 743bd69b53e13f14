"""flatten / unflatten / walk_dict / bounds helpers with the reference's semantics (`diy_gym/utils.py:6-95`),
on batched torch tensors (leading dimension = num_envs) or numpy arrays."""
from collections import OrderedDict

import numpy as np
import torch

from . import spaces


def get_bounds_for_space(space, low_not_high):
    """Per-environment bounds; Dict levels are key-sorted, as the reference does (`utils.py:15-19`)."""
    if isinstance(space, spaces.Discrete):
        return 0 if low_not_high else space.n
    if isinstance(space, spaces.MultiDiscrete):
        return np.zeros(space.nvec.shape) if low_not_high else np.ones(space.nvec.shape) * space.nvec
    if isinstance(space, spaces.MultiBinary):
        return np.zeros(space.n) if low_not_high else np.ones(space.n)
    if isinstance(space, spaces.Box):
        return space.low if low_not_high else space.high
    if isinstance(space, spaces.Dict):
        return OrderedDict((k, get_bounds_for_space(space.spaces[k], low_not_high)) for k in sorted(space.spaces))
    if isinstance(space, spaces.Tuple):
        return tuple(get_bounds_for_space(s, low_not_high) for s in space.spaces)
    try:
        return space.low if low_not_high else space.high
    except AttributeError:
        raise AttributeError('Could not find a bound for the space; custom spaces must define low and high so that they can be flattened')


def get_desc_for_space(space, prepend=''):
    out = []
    for key, val in space.spaces.items():
        if isinstance(val, spaces.Dict):
            out.extend(get_desc_for_space(val, prepend + '/' + key))
        else:
            out.append(prepend + '/' + key)
    return out


def walk_dict(d, func=sum):
    """Collapse a nested dict.  Like the reference (`utils.py:42-43`) only the outermost level
    uses `func`; inner levels are summed.  Leaves may be python scalars or [num_envs] tensors."""
    def inner(e):
        return _tsum(inner(x) if isinstance(x, dict) else x for x in e.values())
    vals = [inner(e) if isinstance(e, dict) else e for e in d.values()]
    if func is sum:
        return _tsum(vals)
    if func is any:
        return _tany(vals)
    if func is all:
        return _tall(vals)
    return func(vals)


def _as_num(v):
    return v.to(torch.float32) if isinstance(v, torch.Tensor) and v.dtype in (torch.bool, torch.uint8) else v


def _tsum(vals):
    tot = 0
    for v in vals:
        tot = tot + _as_num(v)
    return tot


def _tany(vals):
    out = False
    for v in vals:
        out = (v != 0) | out if isinstance(v, torch.Tensor) else (bool(v) | out)
    return out


def _tall(vals):
    out = True
    for v in vals:
        out = (v != 0) & out if isinstance(v, torch.Tensor) else (bool(v) & out)
    return out


def flatten(to_flatten, batched=True):
    """Concatenate every leaf in dict *insertion* order (`utils.py:46-60`).  Batched tensors keep dim 0."""
    leaves = []

    def rec(x):
        if isinstance(x, dict):
            for v in x.values():
                rec(v)
        elif isinstance(x, tuple):
            for v in x:
                rec(v)
        else:
            leaves.append(x)
    rec(to_flatten)
    if any(isinstance(l, torch.Tensor) for l in leaves):
        ref = next(l for l in leaves if isinstance(l, torch.Tensor))
        ts = [l if isinstance(l, torch.Tensor) else torch.as_tensor(l, device=ref.device) for l in leaves]
        if batched:
            return torch.cat([t.reshape(t.shape[0], -1).to(torch.float32) for t in ts], dim=1)
        return torch.cat([t.reshape(-1).to(torch.float32) for t in ts])
    return np.concatenate([np.reshape(l, -1) for l in leaves])


def unflatten(flat, space, batched=True):
    """Inverse of flatten for an action vector laid out with key-sorted Dict levels (`utils.py:63-85`)."""
    pos = [0]

    def take(n):
        pos[0] += n
        return flat[:, pos[0] - n:pos[0]] if batched else flat[pos[0] - n:pos[0]]

    def rec(sp):
        if isinstance(sp, spaces.Dict):
            return OrderedDict((k, rec(sp.spaces[k])) for k in sorted(sp.spaces))
        if isinstance(sp, spaces.Tuple):
            return tuple(rec(s) for s in sp.spaces)
        if isinstance(sp, spaces.Discrete):
            return take(1).round().long()
        if isinstance(sp, spaces.MultiDiscrete):
            return take(int(np.prod(sp.nvec.shape))).round().long()
        if isinstance(sp, spaces.MultiBinary):
            return take(sp.n).round().long()
        if isinstance(sp, spaces.Box):
            n = int(np.prod(sp.shape)) if len(sp.shape) else 1
            v = take(n)
            return v.reshape((v.shape[0], ) + tuple(sp.shape)) if batched else v.reshape(sp.shape)
        raise AttributeError('Unrecognised space in action_space; only the built-in space types can be unflattened')
    return rec(space)
