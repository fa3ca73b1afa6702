"""Generates tests/golden/model_*.json from the reference MJCF assets (run in the build
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_model_golden.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from oracle.mjcf_compile import compile_model  # noqa: E402

ASSETS = '/root/reference/robosumo/robosumo/envs/assets'
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == '__main__':
    for names in (['ant', 'ant'], ['bug', 'bug'], ['spider', 'spider'], ['ant', 'bug'], ['ant', 'spider'], ['bug', 'spider']):
        M = compile_model(ASSETS, names)
        out = os.path.join(HERE, 'model_%s_%s.json' % tuple(names))
        with open(out, 'w') as f:
            json.dump(M, f)
        print(out, M['nq'], M['nv'], M['nbody'], M['ngeom'])
