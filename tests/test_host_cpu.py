"""CPU-side checks of the C++ host layer (generator determinism, batch shapes)."""
import numpy as np

from hypo_b200.hostlib import synth_batch
from tests.oracle_util import oracle_consensus


def test_synth_deterministic_and_thread_independent():
    a = synth_batch(5, 300, 60, 12, "mixed", threads=1)
    b = synth_batch(5, 300, 60, 12, "mixed", threads=4)
    assert a.win.tobytes() == b.win.tobytes() and a.arms.tobytes() == b.arms.tobytes()
    assert a.packed.tobytes() == b.packed.tobytes()
    c = synth_batch(6, 300, 60, 12, "mixed")
    assert c.packed.tobytes() != a.packed.tobytes()


def test_synth_shapes():
    b = synth_batch(1, 50, 120, 30, "internal")
    assert b.n_win == 50 and b.n_arms == 1500
    assert (b.win["n_internal"] == 30).all() and (b.win["n_pre"] == 0).all()
    lens = b.arms["len"]
    assert 100 < lens.mean() < 140
    m = synth_batch(1, 50, 120, 30, "mixed")
    assert (m.win["n_internal"] == 18).all() and (m.win["n_pre"] == 6).all() and (m.win["n_suf"] == 6).all()
    # consensus of clean-ish arms recovers a ~120 bp sequence
    cons, _ = oracle_consensus(b.select(np.arange(4)))
    assert all(100 < len(c) < 140 for c in cons)
