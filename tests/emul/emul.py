"""ctypes wrapper of the CPU emulation of the device step code (tests/emul/emul.cpp).  TEST INFRASTRUCTURE:
lets the GPU-less container check the kernel logic (fp32, team phases) against the fp64 oracle."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, '..', '..', 'diy_gym_b200', 'csrc')
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, 'libdgemul.so')
    srcs = [os.path.join(_HERE, 'emul.cpp')] + [os.path.join(_CSRC, f) for f in ('dg_env.cuh', 'dg_math.cuh', 'dg_scene.h', 'scene_sections.h')]
    if force or not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-std=c++17', '-Wno-unknown-pragmas', '-o', so, srcs[0]])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        fp, ip, dp, vp, u8 = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double),
                              ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint8))
        L.dge_create.restype = vp
        L.dge_create.argtypes = [ip, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.dge_cold_floats.restype = ctypes.c_int
        L.dge_cold_floats.argtypes = [vp]
        L.dge_destroy.argtypes = [vp]
        L.dge_state.restype = fp
        L.dge_state.argtypes = [vp]
        L.dge_param.restype = fp
        L.dge_param.argtypes = [vp]
        L.dge_ws_floats.restype = ctypes.c_int
        L.dge_ws_floats.argtypes = [vp]
        L.dge_set_seed.argtypes = [vp, ctypes.c_uint32, ctypes.c_int]
        L.dge_set_action_mask.argtypes = [vp, u8, ctypes.c_int]
        L.dge_step.argtypes = [vp, fp, fp, fp, u8]
        L.dge_reset.argtypes = [vp, u8, fp, fp, u8]
        L.dge_physics.argtypes = [vp, ctypes.c_int]
        L.dge_observe.argtypes = [vp, fp, fp, u8]
        L.dge_contacts_dropped.restype = ctypes.c_uint
        L.dge_contacts_dropped.argtypes = [vp]
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


class EmulWorld:
    def __init__(self, scene, n_envs=1, team=1, seed=1234, env_off=0, ws_mode=2):
        L = lib()
        self.scene, self.h, self.n = scene, scene.hdr, n_envs
        ib = np.ascontiguousarray(scene.ibuf, np.int32)
        fb = np.ascontiguousarray(scene.fbuf, np.float64)
        self._w = L.dge_create(ib.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ib.size, fb.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                               fb.size, n_envs, team, ws_mode)
        if not self._w:
            raise RuntimeError('emulation rejected the scene')
        S, P = self.h['S'], self.h['P']
        self.state = np.ctypeslib.as_array(L.dge_state(self._w), shape=(n_envs * S + 1, ))[:n_envs * S].reshape(n_envs, S)
        self.param = np.ctypeslib.as_array(L.dge_param(self._w), shape=(n_envs * P + 1, ))[:n_envs * P].reshape(n_envs, P)
        self.ws_floats = L.dge_ws_floats(self._w)
        self.cold_floats = L.dge_cold_floats(self._w)
        L.dge_set_seed(self._w, seed, env_off)

    def __del__(self):
        if getattr(self, '_w', None):
            lib().dge_destroy(self._w)
            self._w = None

    def s(self, name, n, env=0):
        return self.state[env, self.h[name]:self.h[name] + n]

    def p(self, name, n, env=0):
        return self.param[env, self.h[name]:self.h[name] + n]

    def _outptrs(self):
        h = self.h
        self._o = np.zeros((self.n, max(h['n_obs'], 1)), np.float32)
        self._r = np.zeros((self.n, max(h['n_rew'], 1)), np.float32)
        self._t = np.zeros((self.n, max(h['n_term'], 1)), np.uint8)
        if h['n_obs'] == 0 or h['n_rew'] == 0 or h['n_term'] == 0:
            pass  # zero-width rows: the C side never writes them
        return _fp(self._o), _fp(self._r), self._t.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))

    def _outs(self):
        h = self.h
        return self._o[:, :h['n_obs']], self._r[:, :h['n_rew']], self._t[:, :h['n_term']]

    def step(self, act):
        act = np.ascontiguousarray(act, np.float32).reshape(self.n, -1) if self.h['n_act'] else np.zeros((self.n, 1), np.float32)
        o, r, t = self._outptrs()
        lib().dge_step(self._w, _fp(act), o, r, t)
        return self._outs()

    def observe(self):
        o, r, t = self._outptrs()
        lib().dge_observe(self._w, o, r, t)
        return self._outs()

    def contacts_dropped(self):
        return int(lib().dge_contacts_dropped(self._w))

    def reset(self, mask=None):
        o, r, t = self._outptrs()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8).ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        lib().dge_reset(self._w, m, o, r, t)
        return self._outs()

    def physics(self, nsub=None):
        lib().dge_physics(self._w, self.h['substeps'] if nsub is None else nsub)
