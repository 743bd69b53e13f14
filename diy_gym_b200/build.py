"""Builds libdiygym_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc.  Run: python -m diy_gym_b200.build"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.environ.get('DG_LIB') or os.path.join(HERE, 'libdiygym_b200.so')   # DG_LIB: load an alternative build (experiments)
SOURCES = ['dg_kernels.cu']
DEPS = ['dg_kernels.cu', 'dg_env.cuh', 'dg_math.cuh', 'dg_scene.h', 'scene_sections.h', os.path.join('..', '..', 'include', 'diygym_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-shared', '-Xcompiler', '-fPIC']


def stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    build_library(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(LIB)
