#!/usr/bin/env python
"""The four plot families of the reference's results/visualisation_fat_vector.ipynb (cells 9-14: execution time,
speed-up, performance, efficiency — against the number of ranks at a fixed k, and against k at a fixed number of ranks),
regenerated from the CSV tools/sweep.py writes (the reference's get_csv_all.sh columns).

    python tools/plot_results.py profiles/r2_results.csv --matrix cfg2 --out profiles/plots

    speed-up    = Serial Algo Execution time / strategy Execution time     (ipynb cell 10)
    performance = 2 * nnz * k / strategy Execution time, GFLOP/s           (ipynb:1597-1601)
    efficiency  = speed-up / ranks                                         (ipynb cell 14)

Writes plain SVG (no plotting library is installed in this image): one file per family and axis.
"""
from __future__ import annotations

import argparse
import csv
import math
import os

SERIES = [("Serial", "Serial Algo Execution time", "#444444"), ("Row-wise", "Row-wise Execution time", "#1f77b4"),
          ("Column-wise", "Column-wise Execution time", "#d62728"), ("Non-zero elements", "Non-zero Elements Execution time", "#2ca02c")]


def svg_plot(path, title, xlabel, ylabel, xs, curves, logy=False):
    """curves: list of (label, colour, [y or None per x])."""
    W, H, L, R, T, B = 640, 420, 70, 170, 40, 50
    ys = [y for _, _, vals in curves for y in vals if y is not None and (y > 0 or not logy)]
    if not ys or not xs:
        return
    f = (lambda v: math.log10(v)) if logy else (lambda v: v)
    lo, hi = min(map(f, ys)), max(map(f, ys))
    if not logy:
        lo = min(lo, 0.0)
    if hi == lo:
        hi = lo + 1.0
    px = lambda i: L + (W - L - R) * (i / max(1, len(xs) - 1))
    py = lambda v: T + (H - T - B) * (1.0 - (f(v) - lo) / (hi - lo))
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif" font-size="12">',
           f'<rect width="{W}" height="{H}" fill="white"/>', f'<text x="{W / 2}" y="20" text-anchor="middle" font-size="14">{title}</text>',
           f'<line x1="{L}" y1="{H - B}" x2="{W - R}" y2="{H - B}" stroke="black"/>', f'<line x1="{L}" y1="{T}" x2="{L}" y2="{H - B}" stroke="black"/>',
           f'<text x="{(L + W - R) / 2}" y="{H - 12}" text-anchor="middle">{xlabel}</text>',
           f'<text x="16" y="{(T + H - B) / 2}" text-anchor="middle" transform="rotate(-90 16 {(T + H - B) / 2})">{ylabel}</text>']
    for i, x in enumerate(xs):
        out.append(f'<text x="{px(i)}" y="{H - B + 16}" text-anchor="middle">{x}</text>')
    for j in range(5):
        v = lo + (hi - lo) * j / 4
        label = f"{10 ** v:.3g}" if logy else f"{v:.3g}"
        y = T + (H - T - B) * (1 - j / 4)
        out.append(f'<line x1="{L - 4}" y1="{y}" x2="{W - R}" y2="{y}" stroke="#dddddd"/><text x="{L - 8}" y="{y + 4}" text-anchor="end">{label}</text>')
    for n, (label, colour, vals) in enumerate(curves):
        pts = [(px(i), py(v)) for i, v in enumerate(vals) if v is not None and (v > 0 or not logy)]
        if pts:
            out.append(f'<polyline fill="none" stroke="{colour}" stroke-width="2" points="' + " ".join(f"{x:.1f},{y:.1f}" for x, y in pts) + '"/>')
            out += [f'<circle cx="{x:.1f}" cy="{y:.1f}" r="3" fill="{colour}"/>' for x, y in pts]
        out.append(f'<rect x="{W - R + 12}" y="{T + 18 * n}" width="12" height="12" fill="{colour}"/><text x="{W - R + 30}" y="{T + 18 * n + 11}">{label}</text>')
    out.append("</svg>")
    with open(path, "w") as fh:
        fh.write("\n".join(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--matrix", default="cfg2")
    ap.add_argument("--out", default="profiles/plots")
    args = ap.parse_args()
    rows = [r for r in csv.DictReader(open(args.csv)) if r["file Name"].startswith(args.matrix + "_")]
    if not rows:
        raise SystemExit(f"no rows for {args.matrix} in {args.csv}")
    os.makedirs(args.out, exist_ok=True)
    num = lambda r, c: float(r[c]) if r.get(c) not in (None, "", "nan") else None
    k_of = lambda r: int(r["Fat Vector"].split("x")[1])
    ranks = sorted({int(r["Cores Number"]) for r in rows})
    ks = sorted({k_of(r) for r in rows})
    at = {(int(r["Cores Number"]), k_of(r)): r for r in rows}

    def families(xs, key_of, axis_name, fixed_name):
        def col(c):
            return [num(at[key_of(x)], c) if key_of(x) in at else None for x in xs]
        serial = col("Serial Algo Execution time")
        nnz = [float(at[key_of(x)]["nnz"]) if key_of(x) in at else None for x in xs]
        kk = [k_of(at[key_of(x)]) if key_of(x) in at else None for x in xs]
        pp = [int(at[key_of(x)]["Cores Number"]) if key_of(x) in at else None for x in xs]
        times = {label: col(c) for label, c, _ in SERIES}
        tag = f"{args.matrix}_{fixed_name}_vs_{axis_name}".replace(" ", "_").replace("(", "").replace(")", "").replace("=", "")
        svg_plot(os.path.join(args.out, f"{tag}_execution_time.svg"), f"{args.matrix}: execution time ({fixed_name})", axis_name, "seconds", xs,
                 [(label, colour, times[label]) for label, _, colour in SERIES], logy=True)
        div = lambda a, b: [x / y if x and y else None for x, y in zip(a, b)]
        speed = {label: div(serial, times[label]) for label, _, _ in SERIES[1:]}
        svg_plot(os.path.join(args.out, f"{tag}_speedup.svg"), f"{args.matrix}: speed-up over the serial call ({fixed_name})", axis_name, "serial time / time", xs,
                 [(label, colour, speed[label]) for label, _, colour in SERIES[1:]])
        perf = {label: [2.0 * n * k / t / 1e9 if n and k and t else None for n, k, t in zip(nnz, kk, times[label])] for label, _, _ in SERIES}
        svg_plot(os.path.join(args.out, f"{tag}_performance.svg"), f"{args.matrix}: performance ({fixed_name})", axis_name, "GFLOP/s = 2 nnz k / t", xs,
                 [(label, colour, perf[label]) for label, _, colour in SERIES], logy=True)
        eff = {label: [s / p if s and p else None for s, p in zip(speed[label], pp)] for label, _, _ in SERIES[1:]}
        svg_plot(os.path.join(args.out, f"{tag}_efficiency.svg"), f"{args.matrix}: efficiency ({fixed_name})", axis_name, "speed-up / ranks", xs,
                 [(label, colour, eff[label]) for label, _, colour in SERIES[1:]])

    for k in ks:
        if sum((p, k) in at for p in ranks) > 1:
            families(ranks, lambda p, k=k: (p, k), "ranks (GPUs)", f"k={k}")
    for p in ranks:
        if sum((p, k) in at for k in ks) > 1:
            families(ks, lambda k, p=p: (p, k), "k", f"{p} rank{'s' if p > 1 else ''}")
    print(f"wrote {len(os.listdir(args.out))} SVG files to {args.out}")


if __name__ == "__main__":
    main()
