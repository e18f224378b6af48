#!/usr/bin/env python
"""Offline model (CPU, oracle only) of the growth projection the multi-tile tiers use to hand a window on
before it overflows (poa_kernel.cu: add_sequence).  For every window the oracle reports nodes / edges /
DP cells after each sequence; the rule of the kernel is replayed on that curve for a tier's capacities and
compared with what really happens:

  overflow   windows whose DAG really outgrows the tier (nodes or edges)
  caught     ... of those, abandoned by the projection before the overflow; `saved` = share of the DP cells
             a late overflow would have wasted that the early exit avoids
  false      windows that would have fitted but were abandoned (they run in a larger, slower tier)

Usage: python tools/projection_sim.py [--from 6] [--num 3 --den 4] [--margin 8]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from hypo_b200.hostlib import synth_batch  # noqa: E402
from tests.oracle_util import oracle_growth  # noqa: E402

TIERS = {"T0b": (384, 768), "T1m": (640, 1152), "T1": (1024, 1920)}
SHAPES = [  # tier, kwargs
    ("T0b", dict(length=200, n_arms=12, kind="internal", err=0.01)),
    ("T0b", dict(length=125, n_arms=30, kind="internal", err=0.05)),
    ("T0b", dict(length=200, n_arms=30, kind="internal", err=0.02)),
    ("T1m", dict(length=250, n_arms=30, kind="internal", err=0.01)),
    ("T1m", dict(length=250, n_arms=30, kind="internal", err=0.03)),
    ("T1m", dict(length=250, n_arms=30, kind="internal", err=0.05)),
    ("T1m", dict(length=500, n_arms=10, kind="internal", err=0.01)),
    ("T1m", dict(length=500, n_arms=10, kind="internal", err=0.05)),
    ("T1m", dict(length=400, n_arms=20, kind="mixed", err=0.02)),
    ("T1m", dict(length=250, n_arms=30, kind="internal", err=0.01, wtype=1)),
    ("T1m", dict(length=250, n_arms=30, kind="internal", err=0.04, wtype=1)),
    ("T1", dict(length=500, n_arms=30, kind="internal", err=0.01)),
    ("T1", dict(length=500, n_arms=30, kind="internal", err=0.03)),
    ("T1", dict(length=500, n_arms=30, kind="internal", err=0.05)),
    ("T1", dict(length=500, n_arms=30, kind="internal", err=0.02, wtype=1)),
]


def replay(curve, ncap, ecap, a):
    """curve: [n_seq, 3] of one round.  Returns (overflow_at, abandoned_at) - sequence indices or None."""
    n_total = len(curve)
    overflow = abandoned = None
    base = None
    for k in range(n_total):          # k sequences are in the graph when sequence k is about to be added
        nodes, edges = (curve[k - 1][0], curve[k - 1][1]) if k else (0, 0)
        if k == 2:
            base = (nodes, edges)
        elif k >= a.start and abandoned is None and base is not None:
            left, seen = n_total - k, k - 2
            dn, de = nodes - base[0], edges - base[1]
            capn, cape = ncap + ncap // a.margin, ecap + ecap // a.margin
            if a.num * dn * left > a.den * (capn - nodes) * seen or a.num * de * left > a.den * (cape - edges) * seen:
                abandoned = k
        if curve[k][0] > ncap or curve[k][1] > ecap:
            overflow = k
            break
    if abandoned is not None and overflow is not None and abandoned > overflow:
        abandoned = None
    return overflow, abandoned


_CACHE = {}


def curves(kw, n_win):
    key = (str(sorted(kw.items())), n_win)
    if key not in _CACHE:
        b = synth_batch(900 + len(str(kw)), n_win, **kw)
        out = []
        for w in range(b.n_win):
            g = oracle_growth(b, w)
            rounds = [g]
            if kw.get("wtype"):
                d = np.diff(g[:, 0]) < 0
                half = int(np.argmax(d)) + 1 if d.any() else len(g)
                rounds = [g[:half], g[half:]]
            out.append([r for r in rounds if len(r)])
        _CACHE[key] = out
    return _CACHE[key]


def evaluate(a, n_win, verbose=True):
    """Returns (cells wasted by windows that leave the tier, cells of falsely abandoned windows)."""
    tot_waste = tot_false = 0.0
    for tier, kw in SHAPES:
        ncap, ecap = TIERS[tier]
        n_over = n_caught = n_false = 0
        wasted_late = wasted_early = lost = 0.0
        nodes = []
        for rounds in curves(kw, n_win):
            nodes.append(int(max(r[-1][0] for r in rounds)))
            for r in rounds:
                cells = r[:, 2].astype(float)
                ov, ab = replay(r, ncap, ecap, a)
                if ov is not None:
                    late = float(cells[ov - 1] if ov else 0.0)
                    early = float(cells[ab - 1] if ab else 0.0) if ab is not None else late
                    n_over += 1
                    n_caught += ab is not None
                    wasted_late += late
                    wasted_early += early
                    break
                if ab is not None:
                    n_false += 1
                    lost += float(cells[ab - 1]) + a.penalty * float(rounds[-1][-1][2])
                    break
        saved = 1.0 - wasted_early / wasted_late if wasted_late > 0 else 0.0
        tot_waste += wasted_early
        tot_false += lost
        if verbose:
            shape = f"{kw['n_arms']}x{kw['length']} {kw['kind']} err {kw['err']}" + (" LONG" if kw.get("wtype") else "")
            print(f"{tier:4} {shape:40} {int(np.mean(nodes)):6d} {n_over:8d} {n_caught:7d} {100 * saved:5.0f}% {n_false:6d} "
                  f"{lost / 1e6:7.1f}")
    return tot_waste, tot_false


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from", dest="start", type=int, default=6)
    ap.add_argument("--num", type=int, default=3)
    ap.add_argument("--den", type=int, default=4)
    ap.add_argument("--margin", type=int, default=8)
    ap.add_argument("--windows", type=int, default=48)
    ap.add_argument("--penalty", type=float, default=0.5,
                    help="extra cost of a falsely abandoned window, as a share of its DP cells (slower tier)")
    ap.add_argument("--grid", action="store_true", help="evaluate a grid of rule parameters")
    a = ap.parse_args()
    if a.grid:
        print("start num/den margin   wasted by leavers   cost of false alarms   total  (1e6 DP cells)")
        for start in (3, 4, 5, 6, 8):
            for num, den in ((1, 2), (5, 8), (3, 4), (7, 8), (1, 1)):
                for margin in (8, 16, 1000000):
                    b = argparse.Namespace(start=start, num=num, den=den, margin=margin, penalty=a.penalty)
                    w, f = evaluate(b, a.windows, verbose=False)
                    print(f"{start:5d} {num}/{den:<3d} {('1/%d' % margin) if margin < 1000 else '0':>6} {w / 1e6:18.0f} {f / 1e6:22.0f} {(w + f) / 1e6:7.0f}")
        return
    print(f"rule: from sequence {a.start}, {a.num}/{a.den} of the linear extrapolation against capacity * (1 + 1/{a.margin})")
    print(f"{'tier':4} {'shape':40} {'nodes':>6} {'overflow':>8} {'caught':>7} {'saved':>6} {'false':>6} {'lost':>7}")
    evaluate(a, a.windows)


if __name__ == "__main__":
    main()
