#!/usr/bin/env python
"""The "reference on host cores" row: one OS process per core, each building the
reference's own DIYGym (unmodified, `render: no` the only edit) on real pybullet and timing 1 000 step() calls after 100
warm-up steps; aggregate env-steps/s = sum of steps / max wall time.  pybullet is looked for in the environment and under
baseline/_ref (reserved by .gitignore for a driver-supplied install).

When `import pybullet` fails - the case in the build container and on the GPU boxes of round 1 - the script prints
REFERENCE_UNAVAILABLE and exits 0; the CPU row of the result table is then the repo's fp64 oracle, timed by
`bench.py --impl reference` and labelled "oracle (not pybullet)".  Oracle timings are never reported as pybullet timings.

    python baseline/run_pybullet_baseline.py [--config ur_high_5] [--steps 1000] [--warmup 100] [--reference /root/reference]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
CONFIGS = {'ur_high_5': ('examples/ur_high_5/ur_high_5.yaml', 1.0), 'from_the_readme': ('examples/from_the_readme/from_the_readme.yaml', 1.0),
           'drone_pilot': ('examples/drone_pilot/drone_pilot.yaml', 1.0), 'r2d2_maze': (os.path.join(ROOT, 'examples', 'r2d2_maze', 'r2d2_maze.yaml'), 20.0)}


def worker(rank, args, q):
    import numpy as np
    import yaml
    sys.path.insert(0, args.reference)
    sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
    from diy_gym import DIYGym
    from gym import spaces
    path, scale = CONFIGS[args.config]
    node = yaml.load(open(path if os.path.isabs(path) else os.path.join(args.reference, path)), Loader=yaml.FullLoader)
    node['render'] = False
    tmp = os.path.join(tempfile.mkdtemp(), 'cfg.yaml')
    yaml.dump(node, open(tmp, 'w'))
    env = DIYGym(tmp)
    rng = np.random.default_rng(1234 + rank)

    def sample(space):
        if isinstance(space, spaces.Dict):
            return {k: sample(v) for k, v in space.spaces.items()}
        return rng.uniform(np.asarray(space.low, float) * scale, np.asarray(space.high, float) * scale)
    env.reset()
    for _ in range(args.warmup):
        env.step(sample(env.action_space))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, term, _ = env.step(sample(env.action_space))
        if (any(term.values()) if isinstance(term, dict) else term):
            env.reset()
    q.put((rank, args.steps, time.perf_counter() - t0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='r2d2_maze', choices=sorted(CONFIGS))
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--reference', default=os.environ.get('DIYGYM_REFERENCE', '/root/reference'))
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
    try:
        import pybullet  # noqa: F401
        import gym  # noqa: F401
    except ImportError as e:
        print('REFERENCE_UNAVAILABLE: %s - the CPU row is the fp64 oracle (python bench.py --impl reference), labelled "oracle (not pybullet)"' % e)
        return 0
    cores = len(os.sched_getaffinity(0))
    q = mp.Queue()
    procs = [mp.Process(target=worker, args=(r, args, q)) for r in range(cores)]
    for pr in procs:
        pr.start()
    res = [q.get() for _ in procs]
    for pr in procs:
        pr.join()
    total, wall = sum(r[1] for r in res), max(r[2] for r in res)
    print(json.dumps({'impl': 'pybullet DIRECT, reference DIYGym unmodified', 'config': args.config, 'cores': cores, 'value': total / wall,
                      'unit': 'env-steps/s', 'steps_per_process': args.steps, 'warmup': args.warmup}))
    return 0


if __name__ == '__main__':
    sys.exit(main())
