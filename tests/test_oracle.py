"""Pins the CPU restatement (oracle/poa_oracle.c) — the checker every GPU parity test uses —
against (1) the reference's own known-answer test, (2) golden vectors produced by the compiled
unmodified reference, (3) the compiled reference itself when oracle/_ref exists."""
import numpy as np
import pytest

from hypo_b200.batch import WINDOW_LONG, build_batch
from hypo_b200.synth import edge_case_windows, random_batch
from tests.golden_util import group_batch
from tests.oracle_util import (DEFAULT_SCORES, drop_rejected_arms, oracle_consensus, oracle_spoa,
                               ref_consensus, ref_lib, ref_spoa)


def test_spoa_global_consensus_known_answer(golden_spoa):
    """reference external/spoa/test/spoa_test.cpp:220-239 (kNW, 5/-4/-8, 55 reads)."""
    g = golden_spoa
    assert len(g["seqs"]) == 55 and len(g["expected"]) == 469
    assert oracle_spoa(g["seqs"], g["m"], g["n"], g["g"]) == g["expected"]


def test_golden_windows(golden_windows):
    total = 0
    for group in golden_windows["groups"]:
        batch, expected = group_batch(group)
        got, _ = oracle_consensus(batch, tuple(group["scores"]))
        assert got == expected, group["name"]
        total += len(expected)
    assert total >= 200


def test_golden_large_windows(golden_windows_large):
    """Reference-made golden vectors at the sizes of the larger capacity tiers (250-500 bp, 100 reads, LONG)."""
    total = 0
    for group in golden_windows_large["groups"]:
        batch, expected = group_batch(group)
        got, _ = oracle_consensus(batch, tuple(group["scores"]))
        assert got == expected, group["name"]
        total += len(expected)
    assert total == 28


def test_thread_count_independent():
    b = random_batch(11, 64, kind="mixed", length=50, n_arms=12)
    a, _ = oracle_consensus(b, threads=1)
    c, _ = oracle_consensus(b, threads=4)
    assert a == c


def test_gap_must_be_non_positive():
    b = random_batch(12, 2)
    with pytest.raises(RuntimeError):
        oracle_consensus(b, scores=(5, -4, 8, 3, -5, -4))


needs_ref = pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built (no /root/reference)")


@needs_ref
def test_reference_agrees_with_its_own_fixture(golden_spoa):
    g = golden_spoa
    assert ref_spoa(g["seqs"], g["m"], g["n"], g["g"]) == g["expected"]


@needs_ref
@pytest.mark.parametrize("kind", ["internal", "backbone", "prefix", "suffix", "mixed"])
def test_oracle_vs_compiled_reference_short(kind):
    for seed, kw in ((100, dict(length=120, n_arms=30)), (101, dict(length=35, n_arms=14, err=0.08)),
                     (102, dict(length=9, n_arms=25, err=0.03))):
        b = random_batch(seed, 60, kind=kind, **kw)
        r, acc, _ = ref_consensus(b)
        assert acc.all()
        o, _ = oracle_consensus(b)
        assert r == o


@needs_ref
def test_oracle_vs_compiled_reference_long_and_edges():
    for b in (build_batch(edge_case_windows()),
              random_batch(103, 40, kind="mixed", wtype=WINDOW_LONG, length=220, n_arms=14),
              random_batch(104, 30, kind="internal", wtype=WINDOW_LONG, length=400, n_arms=10, err=0.03)):
        r, acc, _ = ref_consensus(b)
        if not acc.all():
            b = drop_rejected_arms(b, acc)
            r, acc, _ = ref_consensus(b)
        o, _ = oracle_consensus(b)
        assert r == o


@needs_ref
def test_oracle_vs_compiled_reference_mixed_alignment_types():
    rng = np.random.default_rng(5)
    from hypo_b200.synth import mutate
    for _ in range(30):
        truth = "".join("ACGT"[i] for i in rng.integers(0, 4, size=60))
        seqs = [mutate(rng, truth, 0.05, 0.05, 0.05) or "A" for _ in range(10)]
        types = [0] + [int(x) for x in rng.integers(0, 3, size=9)]
        assert oracle_spoa(seqs, 5, -4, -8, types) == ref_spoa(seqs, 5, -4, -8, types)


def test_incremental_order_simulation():
    """Between exact sorts the CUDA kernel keeps a valid clique-contiguous topological order
    incrementally (order_update); oracle/poa_oracle.c replays that scheme with
    POA_ORACLE_CHECK_ORDER=1 after every read and aborts if the order is not a valid topological
    order of the graph (edges forward, cliques contiguous)."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from hypo_b200.batch import build_batch, WINDOW_LONG\n"
        "from hypo_b200.synth import random_batch, edge_case_windows\n"
        "from tests.oracle_util import oracle_consensus\n"
        "for b in (build_batch(edge_case_windows()), random_batch(1, 60, kind='mixed'),\n"
        "          random_batch(2, 60, kind='internal', err=0.05), random_batch(3, 40, kind='prefix', length=40, n_arms=20, err=0.1),\n"
        "          random_batch(4, 10, kind='mixed', wtype=WINDOW_LONG, length=250, n_arms=10)):\n"
        "    oracle_consensus(b)\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, POA_ORACLE_CHECK_ORDER="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


def test_growth_trace_is_consistent_with_window_stats():
    """poa_oracle_window_growth (input of tools/projection_sim.py): one entry per added sequence, nodes and
    edges never shrink within a round, and the last entry equals the final graph of poa_oracle_window_stats."""
    from tests.oracle_util import oracle_growth, oracle_stats
    b = random_batch(21, 6, kind="mixed", length=80, n_arms=12, err=0.04)
    for w in range(b.n_win):
        g = oracle_growth(b, w)
        st = oracle_stats(b, w)
        d = b.win[w]
        assert len(g) == int(d["n_internal"]) + int(d["n_pre"]) + int(d["n_suf"]) + (int(d["n_internal"]) == 0)
        assert (np.diff(g[:, 0]) >= 0).all() and (np.diff(g[:, 1]) >= 0).all() and (np.diff(g[:, 2]) > 0).all()
        assert int(g[-1][0]) == int(st[0]) and int(g[-1][1]) == int(st[1]) and int(g[-1][2]) == int(st[2])
    bl = random_batch(22, 3, wtype=WINDOW_LONG, length=150, n_arms=8)
    for w in range(bl.n_win):
        g = oracle_growth(bl, w)
        assert len(g) == 2 * (8 + 1)          # draft / backbone + 8 reads, two rounds
        assert int(g[-1][0]) == int(oracle_stats(bl, w)[0])
