"""pynite_b200: B200-native stiffness assembly + linear solve behind Pynite's `FEModel3D` API.

Import surface mirrors `Pynite/__init__.py:2-12` for the path this package covers.
"""
from pynite_b200.FEModel3D import FEModel3D  # noqa: F401
from pynite_b200.ShearWall import ShearWall    # noqa: F401

__version__ = '0.1.0'
