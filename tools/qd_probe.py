"""Single-step parity probe on the YAML configs: CUDA path vs the fp64 oracle from IDENTICAL states, every step re-synced.

Prints, per config, the worst |dq| / |dqd| against the north_star bar (1e-4 relative, floors as in tests/test_gpu_parity.py),
where the worst entry sits (environment, DoF), whether that environment has contacts in the oracle, and the IK motor targets
(S_MTPOS) of both arms - so that an IK iteration-count flip, an SFU sincos error and a contact-solver difference can be told
apart.  Environment knobs for A/B runs: DG_PRECISE=1 (libm sincos in FK / IK).

    python tools/qd_probe.py [--configs ur_high_5,ur_high_5_randomised] [--envs 64] [--steps 10]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from bench import CONFIGS, action_ranges, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402


def probe(name, n_envs, steps, presteps, verbose, emul=False):
    factory = None
    if emul:   # CPU build of the kernel source (tests/emul): what the GPU-less container can say about the same protocol
        from tests.emul.world import factory as emul_factory
        factory = emul_factory(4)
    env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=n_envs, device=0, seed=4321, world_factory=factory)
    w, sc, h = env.world, env.scene, env.scene.hdr
    nd, nb = sc['nd'], sc['nb']
    lo, hi = action_ranges(env)
    oracles = [OracleWorld(sc, seed=4321, env_id=i) for i in range(n_envs)]
    for o in oracles:
        o.env_reset()
    rng = np.random.default_rng(11)
    for i, o in enumerate(oracles):            # decorrelate the environments
        for _ in range(presteps + i % 4):
            o.env_step(rng.uniform(lo, hi))
    rows = []
    for k in range(steps):
        st = np.stack([o.state for o in oracles]).astype(np.float32)
        w.state.copy_(torch.from_numpy(st))
        w.param.copy_(torch.from_numpy(np.stack([o.param for o in oracles]).astype(np.float32)))
        for i, o in enumerate(oracles):
            o.state[:] = st[i]                 # the oracle starts from the same fp32-rounded row
            o.refresh()
        a = rng.uniform(lo, hi, (n_envs, max(w.n_act, 1))).astype(np.float32)[:, :w.n_act]
        if w.n_act:
            w.action.copy_(torch.from_numpy(a))
        w.step()
        if not emul:
            torch.cuda.synchronize()
        for i, o in enumerate(oracles):
            o.env_step(a[i].astype(np.float64))
        ncon = np.array([len(o.contacts()) for o in oracles])
        sg = w.state.cpu().numpy().astype(np.float64)
        so = np.stack([o.state for o in oracles])
        sl = lambda nm, n: (sg[:, h[nm]:h[nm] + n], so[:, h[nm]:h[nm] + n])
        qg, qo = sl('S_Q', nd)
        vg, vo = sl('S_QD', nd)
        tg, to = sl('S_MTPOS', nd)
        eq, ev, et = np.abs(qg - qo), np.abs(vg - vo), np.abs(tg - to)
        bar_q = 1e-4 * max(np.abs(qo).max(), 1e-30) if nd else 0
        bar_v = 1e-4 * max(np.abs(vo).max(), 1e-2) if nd else 0
        row = {'step': k, 'envs_with_contacts': int((ncon > 0).sum())}
        if nd:
            iv = np.unravel_index(np.argmax(ev), ev.shape)
            row.update({'max_dq': float(eq.max()), 'bar_q': float(bar_q), 'max_dqd': float(ev.max()), 'bar_qd': float(bar_v),
                        'worst_qd_env': int(iv[0]), 'worst_qd_dof': int(iv[1]), 'worst_env_contacts': int(ncon[iv[0]]),
                        'max_dtarget': float(et.max()), 'dtarget_at_worst': float(et[iv]),
                        'max_dqd_contact_free': float(ev[ncon == 0].max()) if (ncon == 0).any() else None,
                        'max_dqd_with_contacts': float(ev[ncon > 0].max()) if (ncon > 0).any() else None})
        for nm, n in (('S_BPOS', 3 * nb), ('S_BQUAT', 4 * nb), ('S_BVEL', 3 * nb), ('S_BOMEGA', 3 * nb)):
            g, o_ = sl(nm, n)
            row['max_d' + nm[2:].lower()] = float(np.abs(g - o_).max())
        rows.append(row)
        if verbose:
            print(json.dumps(row), flush=True)
    env.close()
    worst = {k2: max((r[k2] for r in rows if r.get(k2) is not None), default=None) for k2 in
             ('max_dq', 'max_dqd', 'max_dqd_contact_free', 'max_dqd_with_contacts', 'max_dtarget', 'max_dbpos', 'max_dbquat', 'max_dbvel', 'max_dbomega')}
    return {'config': name, 'envs': n_envs, 'steps': steps, 'precise_sincos': os.environ.get('DG_PRECISE', '0'),
            'bar_q': rows[-1].get('bar_q'), 'bar_qd': rows[-1].get('bar_qd'), 'worst_over_steps': worst, 'per_step': rows}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='ur_high_5,ur_high_5_randomised')
    ap.add_argument('--envs', type=int, default=64)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--presteps', type=int, default=3)
    ap.add_argument('-v', action='store_true')
    ap.add_argument('--emul', action='store_true')
    args = ap.parse_args()
    register_example_addons()
    out = [probe(n, args.envs, args.steps, args.presteps, args.v, args.emul) for n in args.configs.split(',')]
    print(json.dumps(out, indent=1))
