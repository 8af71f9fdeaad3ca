"""Load combination record: name, optional tags, `{case: factor}` (`Pynite/LoadCombo.py:8-21`)."""
from __future__ import annotations


class LoadCombo:
    def __init__(self, name, combo_tags=None, factors=None):
        self.name = name
        self.combo_tags = combo_tags
        self.factors = {} if factors is None else factors

    def __repr__(self):
        return f"LoadCombo(name={self.name!r}, factors={self.factors!r})"

    def AddLoadCase(self, case_name, factor):
        self.factors[case_name] = factor

    def DeleteLoadCase(self, case_name):
        del self.factors[case_name]
