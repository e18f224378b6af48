"""The reference-side binding of INTEGRATION.md §3, compiled against the REFERENCE's own headers and objects
(oracle/ref_binding.cpp -> oracle/_ref/libhypo_ref_binding.so, built in the authoring container where
/root/reference exists; the prebuilt library travels to the GPU box).  Reference hypo::Window objects are
filled through the reference's add_* API, polished on the GPU through the binding's packer and
hypo_gpu_consensus_batch, and then polished again by the reference's own Window::generate_consensus
(reference src/Window.cpp:44-61): the two sets of consensus strings must be identical."""
import ctypes as C
import os

import numpy as np
import pytest

from hypo_b200.batch import WINDOW_LONG, build_batch, split_consensus
from hypo_b200.hostlib import synth_batch
from hypo_b200.synth import edge_case_windows, random_batch

pytestmark = pytest.mark.gpu

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                   "libhypo_ref_binding.so")


def _run(batch, scores=(5, -4, -8, 3, -5, -4)):
    lib = C.CDLL(LIB)
    lib.hypo_refbind_run.restype = C.c_int
    cap = int(batch.out_bound().sum()) + 16
    outs = [np.zeros(cap, np.uint8) for _ in range(2)]
    offs = [np.zeros(batch.n_win + 1, np.uint64) for _ in range(2)]
    sc = (C.c_int8 * 6)(*scores)
    rc = lib.hypo_refbind_run(sc, C.c_int(0), C.c_void_p(batch.win.ctypes.data), C.c_uint64(batch.n_win),
                              C.c_void_p(batch.arms.ctypes.data), C.c_void_p(batch.packed.ctypes.data),
                              C.c_void_p(outs[0].ctypes.data), C.c_void_p(offs[0].ctypes.data),
                              C.c_void_p(outs[1].ctypes.data), C.c_void_p(offs[1].ctypes.data), C.c_uint64(cap))
    assert rc == 0
    return split_consensus(outs[0], offs[0]), split_consensus(outs[1], offs[1])


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libhypo_ref_binding.so was not built (no /root/reference)")
def test_reference_windows_through_the_binding_equal_the_reference_consensus():
    batches = [("edge cases", build_batch(edge_case_windows())),
               ("short mixed", synth_batch(201, 3000, 110, 30, "mixed", 0.03)),
               ("short prefix-heavy", random_batch(202, 96, kind="prefix", length=70, n_arms=24, err=0.05)),
               ("tiny", synth_batch(203, 6000, 9, 30, "mixed", 0.03)),
               ("backbone", random_batch(204, 96, kind="backbone", length=90, n_arms=12, err=0.04)),
               # LONG windows: the reference's Window filters their arms at insert time; both paths see what it kept
               ("long", synth_batch(205, 64, 300, 20, "internal", 0.02, wtype=WINDOW_LONG)),
               ("long noisy", synth_batch(206, 48, 200, 16, "mixed", 0.06, wtype=WINDOW_LONG))]
    for label, b in batches:
        gpu, ref = _run(b)
        bad = [i for i, (x, y) in enumerate(zip(gpu, ref)) if x != y]
        assert not bad, f"{label}: {len(bad)}/{b.n_win} windows differ, first {bad[0]}: {b.spec(bad[0])}"
    gpu, ref = _run(synth_batch(207, 400, 60, 14, "mixed", 0.08), scores=(2, -3, -2, 1, -1, -1))
    assert gpu == ref
