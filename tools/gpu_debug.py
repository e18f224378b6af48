"""Developer aid: run a ladder of batches on the GPU and report mismatches vs the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hypo_b200 import native
from hypo_b200.batch import build_batch, WindowSpec, WINDOW_LONG
from hypo_b200.synth import random_batch, edge_case_windows
from tests.oracle_util import oracle_consensus, DEFAULT_SCORES

def run(name, b, scores=DEFAULT_SCORES):
    native.init(scores, 0)
    t0 = time.time()
    try:
        got = native.consensus(b)
    except Exception as e:
        print(f"[{name}] EXC {e}", flush=True); return
    dt = time.time() - t0
    want, _ = oracle_consensus(b, scores)
    bad = [i for i in range(b.n_win) if got[i] != want[i]]
    print(f"[{name}] n={b.n_win} bad={len(bad)} gpu_time={dt*1e3:.1f}ms launches={native.launch_count()}", flush=True)
    for i in bad[:2]:
        s = b.spec(i)
        print("   win", i, "draft", s.draft[:60], "ni/np/ns", len(s.internal), len(s.pre), len(s.suf), "wt", s.wtype)
        print("   gpu ", got[i][:150]); print("   want", want[i][:150])

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("tiny", "all"):
    run("two-identical", build_batch([WindowSpec("ACGTACGT", ["ACGTACGT", "ACGTACGT"], [], [], 0, 0)]))
    run("sub", build_batch([WindowSpec("ACGTACGT", ["ACGTACGT", "ACGTTCGT", "ACGTACGT"], [], [], 0, 0)]))
    run("indel", build_batch([WindowSpec("ACGTACGT", ["ACGTACGT", "ACGACGT", "ACGTAACGT", "ACGTACGT"], [], [], 0, 0)]))
    run("edge", build_batch(edge_case_windows()))
if which in ("short", "all"):
    for kind in ["internal", "backbone", "prefix", "suffix", "mixed"]:
        run(kind + "-small", random_batch(1, 64, kind=kind, length=30, n_arms=8, err=0.05))
        run(kind + "-30x120", random_batch(2, 64, kind=kind, length=120, n_arms=30))
if which in ("tiers", "all"):
    run("many-nodes", random_batch(8, 12, kind="internal", length=120, n_arms=60, err=0.12))
    run("wide-short", random_batch(9, 8, kind="mixed", length=300, n_arms=10, err=0.03))
    run("long-200", random_batch(5, 24, wtype=WINDOW_LONG, length=200, n_arms=12, kind="mixed"))
    run("long-480", random_batch(6, 16, wtype=WINDOW_LONG, length=480, n_arms=20, kind="internal"))
if which in ("perf", "all"):
    b = random_batch(3, 4096, kind="internal", length=120, n_arms=30)
    run("perf-4096", b)
    run("perf-4096-again", b)
