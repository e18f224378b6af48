#!/usr/bin/env python
"""Window-shape sweep (BASELINE.json configs[4], SURVEY.md §8d "Config 5"): arms x length x read
error, SHORT windows (plus LONG rows), through the host-buffer C ABI on one GPU.

Per shape one JSON line: Mbp/s and GCUPS from the POA kernels' device time, achieved algorithmic
HBM GB/s and its fraction of the measured peak, the tier histogram, and a bit-exact spot check of
the first windows against the CPU oracle (test infrastructure, only used as the checker here).
Run under gpurun:  python tools/sweep.py > gpurun_out/sweep.jsonl
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from hypo_b200 import native  # noqa: E402
from hypo_b200.batch import split_consensus  # noqa: E402
from hypo_b200.hostlib import synth_batch  # noqa: E402
from tests.oracle_util import oracle_consensus, oracle_stats  # noqa: E402

SCORES = (5, -4, -8, 3, -5, -4)


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arms", default="10,30,100,200")
    ap.add_argument("--lengths", default="50,120,250,500")
    ap.add_argument("--errs", default="0.01,0.05")
    ap.add_argument("--budget", type=float, default=2.5e11, help="DP cells per shape (bounds the window count)")
    ap.add_argument("--max-windows", type=int, default=200000)
    ap.add_argument("--long", action="store_true", help="also run LONG windows (lr scores, two rounds) for 30 arms")
    ap.add_argument("--check", type=int, default=512, help="windows compared with the CPU oracle per shape (at most)")
    ap.add_argument("--check-cells", type=float, default=2e10, help="DP cells the oracle may spend per shape")
    ap.add_argument("--only-long", action="store_true", help="LONG rows only")
    ap.add_argument("--long-err", type=float, default=0.01, help="read error of the LONG rows (sub / ins / del each)")
    ap.add_argument("--first-tier", type=int, default=0, help="hypo_gpu_set_option('first_tier'): routing starts there")
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE", help="hypo_gpu_set_option knob (A/B runs)")
    a = ap.parse_args()
    native.init(SCORES, 0)
    for kv in a.option:
        native.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    native.set_option("first_tier", a.first_tier)
    peak = peak_gbs()
    dpx = native.issue_rate(0)   # measured VIADDMNMX.S16x2 rate, 10^9 warp instructions / s
    dpx_peak_gcups = dpx * 32.0  # bench.py: roofline.compute.peak_definition
    shapes = [(int(r), int(l), float(e), 0) for e in a.errs.split(",") for r in a.arms.split(",")
              for l in a.lengths.split(",")]
    if a.only_long:
        shapes = []
    if a.long or a.only_long:
        shapes += [(30, int(l), a.long_err, 1) for l in a.lengths.split(",")]
    for arms, length, err, wtype in shapes:
        probe = synth_batch(99, 4, length, arms, "internal", err, wtype=wtype)
        st = np.array([oracle_stats(probe, w, SCORES) for w in range(2 if arms * length > 30000 else 4)], dtype=np.float64)
        cells = float(st[:, 2].mean()) * (2 if wtype else 1)
        n = int(min(a.max_windows, max(96, a.budget / max(cells, 1.0))))
        batch = synth_batch(1234, n, length, arms, "internal", err, wtype=wtype)
        native.consensus_batch_host(batch)   # warm-up (allocations, first launch)
        t0 = time.perf_counter()
        out, off = native.consensus_batch_host(batch)
        dt = time.perf_counter() - t0
        k_ms, launches, tiers = native.last_timing()
        total = int(off[batch.n_win])
        # windows checked against the oracle: a strided sample over the whole batch, as many as the budget allows
        k = int(max(8, min(a.check, batch.n_win, a.check_cells / max(cells, 1.0))))
        idx = np.unique(np.linspace(0, batch.n_win - 1, k).astype(np.int64))
        allc = split_consensus(out, off)
        got = [allc[i] for i in idx]
        want, _ = oracle_consensus(batch.select(idx), SCORES)
        k = len(idx)
        alg = batch.algorithmic_bytes(total)
        dev_cells = native.last_cells()
        line = {
            "shape": {"arms": arms, "length": length, "err": err, "wtype": "LONG" if wtype else "SHORT"},
            "windows": n, "nodes_avg": float(st[:, 0].mean()), "cells_per_window": cells,
            "kernel_ms": k_ms, "mbp_per_s_kernel": batch.polished_bp / 1e6 / (k_ms / 1e3),
            "mbp_per_s_e2e": batch.polished_bp / 1e6 / dt, "windows_per_s": n / (k_ms / 1e3),
            "gcups": dev_cells / 1e9 / (k_ms / 1e3), "cells_counted_on_device": dev_cells,
            "dpx_peak_gcups": dpx_peak_gcups, "frac_of_dpx_peak": dev_cells / 1e9 / (k_ms / 1e3) / dpx_peak_gcups,
            "rerouted_by_probe": native.last_rerouted(),
            "hbm_gbs_algorithmic": alg / 1e9 / (k_ms / 1e3), "hbm_frac_of_measured_peak": alg / 1e9 / (k_ms / 1e3) / peak,
            "tier_windows": tiers, "abandoned_by_reason": native.last_fail_hist()[1:12],
            "bit_exact_checked": k, "bit_exact": got == want,
        }
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
