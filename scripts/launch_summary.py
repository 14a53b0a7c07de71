"""ncu launch list (``--metrics gpu__time_duration.sum --csv``) -> markdown: the launches of ONE step and per-kernel shares.

usage: python scripts/launch_summary.py launches.csv [--first-kernel k_mol_ae_features] [--bench bench.json] > out.md
"""

from __future__ import annotations

import collections
import csv
import json
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(<[^(]*>)?)", name)
    return m.group(1) if m else name[:60]


def main():
    path = sys.argv[1]
    first = sys.argv[sys.argv.index("--first-kernel") + 1] if "--first-kernel" in sys.argv else "k_mol_ae_features"
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[1:]:
        if r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        launches.append((short(r[ix["Kernel Name"]]), r[ix["Grid Size"]], r[ix["Block Size"]], ms))
    # one step = from the LAST occurrence of `first` to the end of that step (next occurrence or end of list)
    starts = [i for i, l in enumerate(launches) if l[0].startswith(first)]
    if len(starts) >= 2:
        a, b = starts[-2], starts[-1]
    else:
        a, b = (starts[-1] if starts else 0), len(launches)
    step = launches[a:b]
    total = sum(l[3] for l in step)
    print(f"One step = launches {a}..{b - 1} of the capture ({len(step)} launches, {total:.3f} ms under ncu; per-launch times "
          "under ncu are cold-cache and serialised: compare shares, not absolutes).\n")
    print("| # | kernel | grid | block | ms | share |")
    print("|---|---|---|---|---|---|")
    for i, (k, g, bl, ms) in enumerate(step):
        print(f"| {i} | {k} | {g} | {bl} | {ms:.3f} | {100 * ms / total:.1f}% |")
    agg = collections.OrderedDict()
    for k, _, _, ms in step:
        agg[k] = agg.get(k, 0.0) + ms
    print("\n| kernel | ms/step | share |")
    print("|---|---|---|")
    for k, ms in sorted(agg.items(), key=lambda x: -x[1]):
        print(f"| {k} | {ms:.3f} | {100 * ms / total:.1f}% |")
    if "--bench" in sys.argv:
        line = open(sys.argv[sys.argv.index("--bench") + 1]).read().strip().splitlines()[-1]
        json.loads(line)
        print("\nBench line of the same build (not under ncu):\n\n```json\n" + line + "\n```")


if __name__ == "__main__":
    main()
