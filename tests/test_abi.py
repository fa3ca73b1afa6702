"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/rs_b200.h declares."""
import ctypes
import os
import re

from robosumo_selfplay_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    build.build()
    L = _lib.lib()
    header = open(os.path.join(ROOT, 'include', 'rs_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(rs_[a-z_0-9]+)\s*\(', header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), "missing export: %s" % name
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    assert L.rs_agent_model_size() == ctypes.sizeof(_lib.rs_agent_model)


def test_product_has_no_oracle_import():
    """The product package must not reach into oracle/ (or the test-only host emulation)."""
    pkg = os.path.join(ROOT, 'robosumo_selfplay_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'libemu' not in src, f
