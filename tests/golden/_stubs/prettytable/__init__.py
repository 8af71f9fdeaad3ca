"""Throw-away stand-in for the `prettytable` package (absent in this image).

Only needed so that `import Pynite` (the read-only reference under /root/reference) succeeds when
tests/golden/make_golden.py generates fixtures; nothing on the assembly/solve path uses it.
"""


class PrettyTable:
    def __init__(self):
        self.field_names = []
        self._rows = []

    def add_row(self, row):
        self._rows.append(list(row))

    def __str__(self):
        lines = [' | '.join(str(c) for c in self.field_names)]
        lines += [' | '.join(str(c) for c in r) for r in self._rows]
        return '\n'.join(lines)
