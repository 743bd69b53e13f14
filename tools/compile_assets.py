#!/usr/bin/env python
"""Compile the reference's vendored URDFs into model descriptors (JSON) committed under
`diy_gym_b200/data/compiled/`, so the GPU box needs neither the 21 MB mesh tree nor /root/reference.

Usage:  python tools/compile_assets.py [/root/reference/diy_gym/data]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from diy_gym_b200.compiler.urdf import compile_urdf, save_model  # noqa: E402

WANTED = ['ur5/ur5_robot.urdf', 'jaco/j2s7s300_standalone.urdf', 'hector_quadrotor/quadrotor.urdf', 'grass/plane.urdf',
          'plain_plane/plane.urdf', 'wall/wall.urdf', 'ur5/ur5_2f.urdf', 'ur5/ur5_3f.urdf', 'robotiq_2f/gripper.urdf', 'robotiq_3f/gripper.urdf']


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/diy_gym/data'
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'diy_gym_b200', 'data', 'compiled')
    for rel in WANTED:
        desc = compile_urdf(os.path.join(src, rel), rel_name=rel)
        out = os.path.join(dst, rel + '.json')
        os.makedirs(os.path.dirname(out), exist_ok=True)
        save_model(desc, out)
        nd = sum(1 for l in desc['links'] if l['joint'] and l['joint']['type'] != 'fixed')
        print('%-40s links=%2d dof=%2d -> %s (%d B)' % (rel, len(desc['links']), nd, os.path.relpath(out), os.path.getsize(out)))


if __name__ == '__main__':
    main()
