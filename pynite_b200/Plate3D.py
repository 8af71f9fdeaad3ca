"""Import-surface alias: the reference exposes `Pynite.Plate3D.Plate3D` (`Plate3D.py:14`)."""
from pynite_b200.Quad3D import Plate3D  # noqa: F401
