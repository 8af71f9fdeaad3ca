"""pytest plugin (`-p tests.pynite_alias`) that lets the REFERENCE'S OWN test files run against this package: `import Pynite`
resolves to `pynite_b200`, and -- there being no GPU in the development container -- the engine is the oracle-backed
stand-in of tests/oracle_engine.py.  Test infrastructure for tests/test_reference_suite.py; development container only
(the reference's tests live in /root/reference/Testing, which does not travel)."""
import importlib
import sys
import types

import pynite_b200
from pynite_b200 import analysis, engine
from tests import oracle_engine

sys.modules['Pynite'] = pynite_b200
for name in ('FEModel3D', 'Node3D', 'Member3D', 'PhysMember', 'Quad3D', 'Plate3D', 'Spring3D', 'Mesh', 'LoadCombo', 'Material',
             'Section', 'ShearWall', 'MatFoundation'):
    module = importlib.import_module('pynite_b200.' + name)
    sys.modules['Pynite.' + name] = module
    setattr(pynite_b200, name, getattr(pynite_b200, name, module))


# `from Pynite import Analysis` (the reference's tests reach into it for one private helper; the module here is the drivers')
sys.modules['Pynite.Analysis'] = analysis
pynite_b200.Analysis = analysis


def _engine_of(model):
    if not isinstance(getattr(model, '_engine', None), oracle_engine.OracleEngine):
        model._engine = oracle_engine.OracleEngine()
    return model._engine


analysis._engine = _engine_of
engine.Engine = type('Engine', (oracle_engine.OracleEngine,), {'make_opts': staticmethod(engine.Engine.make_opts)})


class _NoDisplay:
    """Stands in for the renderers some of the reference's tests construct after their assertions."""
    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        return lambda *args, **kwargs: None


for name in ('Rendering', 'Visualization'):
    stub = types.ModuleType('Pynite.' + name)
    stub.Renderer = _NoDisplay
    stub.render_model = lambda *args, **kwargs: None
    sys.modules['Pynite.' + name] = stub
