"""Model: one spawned URDF (what `diy_gym/model.py:11-106` of the reference builds through pybullet).

At construction a Model registers its body with the scene builder (north_star (a): URDF -> SoA buffers) and
creates its add-ons and nested models; at run time it is a thin view on the batched device state for user
add-ons: every query returns torch tensors with leading dimension `num_envs`.

Config keys, as in the reference (`model.py:34-47`): model, xyz, rpy, scale, use_fixed_base, mass, color,
parent_frame, child_frame.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import torch_math as tm
from .addons.addon import AddonFactory, Receptor
from .assets import resolve_model
from .compiler.mathutil import quat_from_euler


class Model(Receptor):
    def __init__(self, config, parent=None, env=None):
        Receptor.__init__(self)
        self.name = config.name
        self.env = env if env is not None else parent.env
        self.position = config.get('xyz', [0., 0., 0.])
        self.orientation = quat_from_euler(config.get('rpy', [0., 0., 0.]))
        use_fixed_base = config.get('use_fixed_base', False)
        scale = config.get('scale', 1.0)
        urdf = config.get('model')
        desc = resolve_model(urdf)   # raises ValueError('Could not find URDF: ...') like model.py:63
        spawn_pos, spawn_quat = self.position, self.orientation
        if parent is not None:
            # model.py:69-77: the child is spawned at the COM pose of the parent frame (getLinkState[:2] / base pose) and welded
            # to it with p.createConstraint(JOINT_FIXED): joint frame = (xyz, rpy) in the parent frame, identity in the child frame
            parent_frame_id = parent.get_frame_id(config.get('parent_frame')) if 'parent_frame' in config else -1
            # (the parent's controllers have already put its joints at their rest_position: add-ons are built before nested models)
            # - but only the controllers whose constructor calls reset() in the reference do that (ik_controller.py:45,
            # admittance_controller.py:34); under a joint_controller the parent is still at q = 0 when the child is welded
            q_rest = {}
            for a in parent.addons.values():
                if getattr(a, 'resets_in_constructor', False) and hasattr(a, 'joint_ids') and hasattr(a, 'rest_position'):
                    q_rest.update(dict(zip(a.joint_ids, a.rest_position)))
            T = parent.body.rest_com_pose(parent_frame_id, q_rest)
            spawn_pos, spawn_quat = T.p, T.q
        self.body = self.env.builder.add_body(self.name, desc, xyz=spawn_pos, quat=spawn_quat, scale=scale,
                                              fixed_base=use_fixed_base, mass=config.get('mass') if 'mass' in config else None,
                                              color=config.get('color') if 'color' in config else None)
        if parent is not None:
            child_frame_id = self.get_frame_id(config.get('child_frame')) if 'child_frame' in config else -1
            self.env.builder.add_fixed_constraint(parent.body, parent_frame_id, self.body, child_frame_id, self.position, self.orientation)
        self.uid = self.body.index
        self.addons = OrderedDict(sorted({child.name: AddonFactory.build(child.get('addon'), self, child)
                                          for child in config.find_all('addon')}.items(), key=lambda t: t[0]))
        self.models = OrderedDict(sorted({child.name: Model(child, self) for child in config.find_all('model')}.items(),
                                         key=lambda t: t[0]))

    # ---- pybullet-like introspection (compile time) -----------------------------------------------
    def get_frame_id(self, frame):
        """Joint index of the named joint / frame, -1 for the base or an unknown name (`model.py:94-96`)."""
        return self.body.joint_index(frame)

    def num_joints(self):
        return self.body.num_joints()

    def joint_info(self, i):
        return self.body.joint_info(i)

    # ---- batched state views (run time) --------------------------------------------------------
    @property
    def world(self):
        return self.env.world

    def _h(self, name):
        return self.env.scene.hdr[name]

    def base_pose(self):
        """COM pose of the base: ([N,3], [N,4] xyzw) - getBasePositionAndOrientation."""
        b, st = self.uid, self.world.state
        return st[:, self._h('S_BPOS') + 3 * b:self._h('S_BPOS') + 3 * b + 3], st[:, self._h('S_BQUAT') + 4 * b:self._h('S_BQUAT') + 4 * b + 4]

    def base_velocity(self):
        b, st = self.uid, self.world.state
        return st[:, self._h('S_BVEL') + 3 * b:self._h('S_BVEL') + 3 * b + 3], st[:, self._h('S_BOMEGA') + 3 * b:self._h('S_BOMEGA') + 3 * b + 3]

    def link_state(self, frame_id):
        """COM-frame pose and world velocity of link `frame_id` (-1 = base) as cached after the last step:
        (pos, quat, lin_vel, ang_vel) - getLinkState[0,1,6,7] with computeLinkVelocity=1."""
        if frame_id < 0:
            return self.base_pose() + self.base_velocity()
        gl, st = self.body.link_start + frame_id, self.world.state
        return (st[:, self._h('S_LPOS') + 3 * gl:self._h('S_LPOS') + 3 * gl + 3], st[:, self._h('S_LQUAT') + 4 * gl:self._h('S_LQUAT') + 4 * gl + 4],
                st[:, self._h('S_LVEL') + 3 * gl:self._h('S_LVEL') + 3 * gl + 3], st[:, self._h('S_LOMEGA') + 3 * gl:self._h('S_LOMEGA') + 3 * gl + 3])

    def get_transform(self, frame_id=-1):
        """URDF link-frame pose (getLinkState[4,5]) or the base pose (`model.py:98-106`)."""
        if frame_id < 0:
            return self.base_pose()
        pos, quat, _, _ = self.link_state(frame_id)
        lf = self.env.scene.sec['LINK_F'][self.body.link_start + frame_id]
        d = pos.new_tensor(lf[7:10])
        qi = quat.new_tensor([-lf[16], -lf[17], -lf[18], lf[19]])
        return pos - tm.quat_rotate(quat, d.expand_as(pos)), tm.quat_mul(quat, qi.expand_as(quat))

    def joint_state(self, joint_ids=None):
        """(q, qd, applied motor torque) for the given joint indices (default: all movable) - getJointStates[0,1,3]."""
        ids = self.body.movable_joints() if joint_ids is None else list(joint_ids)
        dofs = [self.body.global_dof(i) for i in ids]
        st, dt = self.world.state, self.env.scene.hdr_f['dt']
        idx = torch.as_tensor(dofs, device=st.device, dtype=torch.long)
        return st[:, self._h('S_Q') + idx], st[:, self._h('S_QD') + idx], st[:, self._h('S_MAPPLIED') + idx] / dt

    def _frame(self, link):
        return self.body.frame(link)

    def apply_external_force(self, link, force, position=None, frame='world'):
        """Batched applyExternalForce: force [N,3] (or [3]) on link (-1 = base) at `position`, both given in
        the WORLD frame or in the LINK (COM) frame; acts during the next step only."""
        st = self.world.state
        f = self._frame(link)
        force = torch.as_tensor(force, device=st.device, dtype=torch.float32).expand(st.shape[0], 3)
        pos, quat, _, _ = self.link_state(link)
        if position is None:
            position = pos if frame == 'world' else torch.zeros_like(pos)
        position = torch.as_tensor(position, device=st.device, dtype=torch.float32).expand(st.shape[0], 3)
        if frame == 'link':
            force = tm.quat_rotate(quat, force)
            rel = tm.quat_rotate(quat, position)
        else:
            rel = position - pos
        o_f, o_t = self._h('S_EXTF') + 3 * f, self._h('S_EXTT') + 3 * f
        st[:, o_f:o_f + 3] += force
        st[:, o_t:o_t + 3] += torch.cross(rel, force, dim=-1)

    def apply_external_torque(self, link, torque, frame='world'):
        st = self.world.state
        f = self._frame(link)
        torque = torch.as_tensor(torque, device=st.device, dtype=torch.float32).expand(st.shape[0], 3)
        if frame == 'link':
            torque = tm.quat_rotate(self.link_state(link)[1], torque)
        o_t = self._h('S_EXTT') + 3 * f
        st[:, o_t:o_t + 3] += torque
