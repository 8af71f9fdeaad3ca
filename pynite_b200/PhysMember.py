"""Import-surface alias: the reference exposes `Pynite.PhysMember.PhysMember` (`PhysMember.py:17`)."""
from pynite_b200.Member3D import PhysMember  # noqa: F401
