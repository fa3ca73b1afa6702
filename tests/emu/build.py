"""Compiles the kernel source as a host emulation (tests/_emu/libemu.so) -- test harness only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, 'tests', '_emu', 'libemu.so')
DEPS = [os.path.join(HERE, 'emu.cpp'), os.path.join(ROOT, 'robosumo_selfplay_b200', 'csrc', 'rs_core.h'),
        os.path.join(ROOT, 'robosumo_selfplay_b200', 'csrc', 'rs_env.h'), os.path.join(ROOT, 'include', 'rs_b200.h')]


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-o', OUT, DEPS[0], '-lm'])
    return OUT
