"""Summarise an ``ncu --set full --import-source on`` report: headline metrics per captured launch and the SASS
instructions that collect the most warp-stall samples (with the dominant stall reason).

usage: python scripts/ncu_summary.py report.ncu-rep [--top 25] [--md out.md]
"""

from __future__ import annotations

import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "derived__lts__lts2xbar_bytes.sum.per_second",
    "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_l1tex2xbar_write_bytes.sum",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "smsp__cycles_active.avg",
]


def ncu(args):
    return subprocess.run(["ncu", *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    out = []
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, body = raw[0], raw[1], raw[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    # every tensor-pipe counter the report holds (names differ between ncu versions / metric sets)
    for h in hdr:
        if ("pipe_tensor" in h or "pipe_tc" in h or "tcgen" in h.lower()) and h not in METRICS:
            METRICS.append(h)
    out.append("| launch | " + " | ".join(m for m in METRICS if m in ix) + " |")
    out.append("|---|" + "---|" * sum(m in ix for m in METRICS))
    for r in body:
        name = r[ix["Kernel Name"]][:40]
        vals = [f"{r[ix[m]]} {units[ix[m]]}" for m in METRICS if m in ix]
        out.append(f"| {r[ix['ID']]} {name} | " + " | ".join(vals) + " |")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    for si, s in enumerate(secs):
        e = secs[si + 1] if si + 1 < len(secs) else len(src)
        h = src[s + 1]
        jx = {k: i for i, k in enumerate(h)}
        rows = [r for r in src[s + 2:e] if len(r) > 5]
        total = sum(int(r[jx["# Samples"]]) for r in rows)
        out.append(f"\nlaunch {si}: {src[s][1][:60]} -- {total} stall samples, top instructions")
        out.append("| samples | % | address | SASS | dominant stall |")
        out.append("|---|---|---|---|---|")
        stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        for r in sorted(rows, key=lambda r: -int(r[jx["# Samples"]]))[:top]:
            st = sorted(((int(r[jx[k]] or 0), k[6:]) for k in stall_cols), reverse=True)[:2]
            dom = ", ".join(f"{k} {v}" for v, k in st if v)
            n = int(r[jx["# Samples"]])
            out.append(f"| {n} | {100.0 * n / max(total, 1):.1f} | {r[jx['Address']][-5:]} | `{r[jx['Source']].strip()[:70]}` | {dom} |")
    if "--json" in sys.argv:
        import json
        def num(r, m):
            v = float(r[ix[m]].replace(",", ""))
            u = units[ix[m]].lower()
            return v * (1e9 if u.startswith("g") else 1e6 if u.startswith("m") else 1e3 if u.startswith("k") else 1)
        js = [{"kernel": r[ix["Kernel Name"]][:60], "duration_ms": float(r[ix["gpu__time_duration.sum"]]) * (1e-3 if units[ix["gpu__time_duration.sum"]].startswith("u") else 1),
               "dram_bytes": num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")} for r in body]
        open(sys.argv[sys.argv.index("--json") + 1], "w").write(json.dumps(js, indent=1) + "\n")
    text = "\n".join(out)
    if "--md" in sys.argv:
        open(sys.argv[sys.argv.index("--md") + 1], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
