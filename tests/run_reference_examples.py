"""Helper of tests/test_reference_suite.py: runs the reference's example scripts (paths on the command line) in this
process with `Pynite` aliased to `pynite_b200` (tests/pynite_alias.py), the oracle-backed engine stand-in, and
do-nothing stand-ins for matplotlib and the renderers.  Prints one `OK <name>` / `FAIL <name> <error>` line per script."""
import contextlib
import io
import os
import runpy
import sys
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.figure', 'matplotlib.axes'):
    sys.modules[name] = MagicMock()
sys.modules['matplotlib.pyplot'].subplots = lambda **k: (MagicMock(), MagicMock())
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']

import tests.pynite_alias  # noqa: E402,F401
import pynite_b200         # noqa: E402

pynite_b200.Visualization = sys.modules['Pynite.Visualization']
pynite_b200.Rendering = sys.modules['Pynite.Rendering']

for path in sys.argv[1:]:
    name = os.path.basename(path)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            runpy.run_path(path, run_name='__main__')
        print('OK', name)
    except BaseException as err:                # noqa: BLE001
        print('FAIL', name, type(err).__name__, str(err)[:200])
