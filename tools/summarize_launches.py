"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline='') as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
header = next(rd)
ki, vi, ui = header.index('Kernel Name'), header.index('Metric Value'), header.index('Metric Unit')
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if len(r) <= vi:
        continue
    name = re.sub(r'\(.*', '', r[ki])
    name = re.sub(r'^void ', '', name)
    v = float(r[vi].replace(',', ''))
    unit = r[ui]
    ns = v*{'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print(f'# ncu launch list summary: {path}')
print(f'\n{sum(cnt.values())} launches, {total/1e6:.2f} ms summed device time (cold-cache, serialised: compare shares)\n')
print('| kernel | launches | total ms | share | avg us |')
print('|---|---:|---:|---:|---:|')
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f'| `{k[:90]}` | {cnt[k]} | {tot[k]/1e6:.3f} | {100*tot[k]/total:.1f}% | {tot[k]/cnt[k]/1e3:.1f} |')
