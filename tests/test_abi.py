"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol the header declares,
and the product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as ct
import os
import re

import pytest

from pynite_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'pynite_b200.h')).read()
    return sorted(set(re.findall(r'PN_EXPORT\s+[\w\s\*]+?\b(pn_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(engine.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = ct.CDLL(engine.library_path())
    names = _header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    # the ctypes prototype table covers exactly the header
    assert sorted(engine.SIGNATURES.keys()) == names


def test_version_string():
    lib = engine.load_library()
    assert b'sm_100a' in lib.pn_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from pynite_b200 import FEModel3D
    from tests import models
    m = models.space_frame(FEModel3D)
    with pytest.raises(engine.EngineUnavailable):
        m.analyze()
    with pytest.raises(engine.EngineUnavailable):
        m.analyze_linear(sparse=False)         # the dense flag does not select a CPU solver either


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, 'pynite_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cpp')):
                src = open(os.path.join(dirpath, f), errors='ignore').read()
                assert 'import oracle' not in src and 'from oracle' not in src, f
