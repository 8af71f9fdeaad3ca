"""Import-surface alias: the reference exposes `Pynite.Section.Section` (`Pynite/Section.py:11`)."""
from pynite_b200.Material import Section, SteelSection  # noqa: F401
