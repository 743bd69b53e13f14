"""ctypes wrapper of the CPU oracle (oracle/bullet_restatement.c).  TEST INFRASTRUCTURE - see that file's header.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, 'liboracle.so')
    srcs = [os.path.join(_HERE, f) for f in ('bullet_restatement.c', 'oracle_batch.c', 'scene_sections.h', 'Makefile')]
    if force or not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'liboracle.so')
        if not os.path.isfile(so):
            build()
        L = ctypes.CDLL(so)
        dp, ip, vp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p
        L.dgo_create.restype = vp
        L.dgo_create.argtypes = [ip, ctypes.c_int, dp, ctypes.c_int]
        L.dgo_destroy.argtypes = [vp]
        L.dgo_state.restype = dp
        L.dgo_state.argtypes = [vp]
        L.dgo_params.restype = dp
        L.dgo_params.argtypes = [vp]
        for f in ('dgo_state_size', 'dgo_param_size', 'dgo_num_contacts'):
            getattr(L, f).restype = ctypes.c_int
            getattr(L, f).argtypes = [vp]
        L.dgo_set_seed.argtypes = [vp, ctypes.c_uint32, ctypes.c_int]
        L.dgo_contacts_dropped.restype = ctypes.c_long
        L.dgo_contacts_dropped.argtypes = [vp]
        L.dgo_flops.restype = ctypes.c_double
        L.dgo_flops.argtypes = [ctypes.c_int]
        L.dgo_get_contact.argtypes = [vp, ctypes.c_int, dp]
        for f in ('dgo_forward_kinematics', 'dgo_refresh', 'dgo_step_physics', 'dgo_env_reset'):
            getattr(L, f).argtypes = [vp]
        L.dgo_frame_state.argtypes = [vp, ctypes.c_int, dp]
        L.dgo_ik.argtypes = [vp, ctypes.c_int, ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp]
        L.dgo_apply_actions.argtypes = [vp, dp]
        u8 = ctypes.POINTER(ctypes.c_uint8)
        L.dgo_observe.argtypes = [vp, dp, dp, u8]
        L.dgo_env_step.argtypes = [vp, dp, dp, dp, u8]
        L.dgo_render.argtypes = [vp, ctypes.c_int, dp, dp]
        L.dgo_render_seg.argtypes = [vp, ctypes.c_int, dp, dp, dp]
        L.dgo_get_camera_image.argtypes = [vp, ctypes.c_int, ctypes.c_int, dp, dp, ctypes.POINTER(ctypes.c_uint8), dp, ctypes.POINTER(ctypes.c_int32)]
        L.dgo_batch_max_threads.restype = ctypes.c_int
        L.dgo_batch_step.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, dp, ctypes.c_int, dp, ctypes.c_int, dp,
                                     ctypes.c_int, u8, ctypes.c_int, ctypes.c_int]
        L.dgo_batch_reset.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class OracleWorld:
    """One fp64 environment built from a compiled `Scene`."""
    def __init__(self, scene, seed=1234, env_id=0):
        L = lib()
        self.scene = scene
        self.h = scene.hdr
        ib = np.ascontiguousarray(scene.ibuf, np.int32)
        fb = np.ascontiguousarray(scene.fbuf, np.float64)
        self._w = L.dgo_create(ib.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ib.size, _dp(fb), fb.size)
        if not self._w:
            raise RuntimeError('oracle rejected the scene buffers')
        self.state = np.ctypeslib.as_array(L.dgo_state(self._w), shape=(max(self.h['S'], 1), ))[:self.h['S']]
        self.param = np.ctypeslib.as_array(L.dgo_params(self._w), shape=(max(self.h['P'], 1), ))[:self.h['P']]
        L.dgo_set_seed(self._w, seed, env_id)
        self.obs = np.zeros(max(self.h['n_obs'], 1))
        self.rew = np.zeros(max(self.h['n_rew'], 1))
        self.term = np.zeros(max(self.h['n_term'], 1), np.uint8)

    def __del__(self):
        if getattr(self, '_w', None):
            lib().dgo_destroy(self._w)
            self._w = None

    def s(self, name, n):
        return self.state[self.h[name]:self.h[name] + n]

    def p(self, name, n):
        return self.param[self.h[name]:self.h[name] + n]

    def refresh(self):
        lib().dgo_refresh(self._w)

    def step_physics(self):
        lib().dgo_step_physics(self._w)

    def apply_actions(self, act):
        act = np.ascontiguousarray(act, np.float64)
        lib().dgo_apply_actions(self._w, _dp(act))

    def observe(self):
        lib().dgo_observe(self._w, _dp(self.obs), _dp(self.rew), self.term.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return self.obs[:self.h['n_obs']].copy(), self.rew[:self.h['n_rew']].copy(), self.term[:self.h['n_term']].copy()

    def env_step(self, act):
        act = np.ascontiguousarray(act, np.float64)
        if act.size == 0:
            act = np.zeros(1)
        lib().dgo_env_step(self._w, _dp(act), _dp(self.obs), _dp(self.rew), self.term.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
        return self.obs[:self.h['n_obs']].copy(), self.rew[:self.h['n_rew']].copy(), self.term[:self.h['n_term']].copy()

    def env_reset(self):
        lib().dgo_env_reset(self._w)
        return self.observe()

    def frame_state(self, frame):
        out = np.zeros(20)
        lib().dgo_frame_state(self._w, frame, _dp(out))
        return dict(com_pos=out[0:3], com_quat=out[3:7], vel=out[7:10], omega=out[10:13], link_pos=out[13:16], link_quat=out[16:20])

    def ik(self, body, ee_link_global, tpos, torn=None, nullspace=None):
        nd = int(self.scene.sec['BODY_I'][body][4])
        out = np.zeros(max(nd, 1))
        tpos = np.ascontiguousarray(tpos, np.float64)
        tq = np.ascontiguousarray(torn if torn is not None else [0, 0, 0, 1], np.float64)
        z = np.zeros(max(nd, 1))
        lo, hi, rng, rest = [np.ascontiguousarray(a, np.float64) for a in (nullspace if nullspace is not None else (z, z, z, z))]
        lib().dgo_ik(self._w, body, ee_link_global, _dp(tpos), _dp(tq), int(torn is not None), int(nullspace is not None), _dp(lo),
                     _dp(hi), _dp(rng), _dp(rest), _dp(out))
        return out[:nd]

    def contacts_dropped(self):
        return int(lib().dgo_contacts_dropped(self._w))

    def contacts(self):
        n = lib().dgo_num_contacts(self._w)
        out = []
        for i in range(n):
            c = np.zeros(13)
            lib().dgo_get_contact(self._w, i, _dp(c))
            out.append(dict(fa=int(c[0]), fb=int(c[1]), pa=c[2:5], pb=c[5:8], n=c[8:11], dist=c[11], mu=c[12]))
        return out

    def get_camera_image(self, width, height, view, proj):
        """p.getCameraImage: camera given by column-major view / projection matrices only -> (rgba u8 [H,W,4], depth buffer [H,W], ids [H,W])."""
        rgba = np.zeros((height, width, 4), np.uint8)
        depth = np.zeros((height, width))
        segm = np.zeros((height, width), np.int32)
        v, pm = np.ascontiguousarray(view, np.float64).reshape(-1), np.ascontiguousarray(proj, np.float64).reshape(-1)
        lib().dgo_get_camera_image(self._w, int(width), int(height), _dp(v), _dp(pm), rgba.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), _dp(depth),
                                   segm.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        return rgba, depth, segm

    def render(self, cam=0, seg=False):
        w, hgt = int(self.scene.sec['CAM_I'][cam][1]), int(self.scene.sec['CAM_I'][cam][2])
        rgb = np.zeros((hgt, w, 3))
        depth = np.zeros((hgt, w))
        if seg:
            mask = np.zeros((hgt, w))
            lib().dgo_render_seg(self._w, cam, _dp(rgb), _dp(depth), _dp(mask))
            return rgb, depth, mask
        lib().dgo_render(self._w, cam, _dp(rgb), _dp(depth))
        return rgb, depth


def flops(reset=False):
    return lib().dgo_flops(int(reset))
