"""Builds robosumo_selfplay_b200/librs_b200.so (hand-written sm_100a CUDA + the C ABI of include/rs_b200.h).

    python -m robosumo_selfplay_b200.build
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'rs_api.cu')
OUT = os.path.join(HERE, 'librs_b200.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in ('rs_api.cu', 'rs_core.h', 'rs_env.h', 'rs_learn.cuh', 'rs_learn_tc.cuh', 'rs_tc.cuh')] + \
       [os.path.join(os.path.dirname(HERE), 'include', 'rs_b200.h')]


def nvcc():
    for p in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return 'nvcc'


def build(force=False, verbose=False):
    if not force and not os.environ.get('RS_DEV_ANT_ONLY') and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = [nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC', '-o', OUT, SRC]
    if os.environ.get('RS_DEV_ANT_ONLY'):      # developer build: Ant-vs-Ant only, written next to the variants; select it with RS_B200_LIB
        cmd.insert(1, '-DRS_DEV_ANT_ONLY')
        dev = os.path.join(os.path.dirname(HERE), 'build', 'variants', 'librs_dev.so')
        os.makedirs(os.path.dirname(dev), exist_ok=True)
        cmd[cmd.index(OUT)] = dev
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
    print(OUT)
