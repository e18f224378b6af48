"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/hypo_b200.h declares, the descriptor layouts match, and the product path fails loudly
(no CPU fallback) when no device is usable."""
import ctypes
import os
import re

import numpy as np
import pytest

from hypo_b200 import native
from hypo_b200.batch import ARM_DTYPE, WIN_DTYPE, build_batch, pack2, pack4, unpack2, unpack4, WindowSpec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "hypo_b200.h")).read()
    return sorted(set(re.findall(r"\b(hypo_gpu_\w+)\s*\(", hdr)))


def test_header_symbols_exported():
    if not os.path.exists(native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(native.LIB_PATH)
    declared = _declared_symbols()
    assert set(declared) == set(native.ABI_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.hypo_gpu_abi_version.restype = ctypes.c_int
    assert lib.hypo_gpu_abi_version() == 2


def test_descriptor_layouts_match_header():
    assert WIN_DTYPE.itemsize == 40 and ARM_DTYPE.itemsize == 16
    assert [WIN_DTYPE.fields[n][1] for n in ("draft_off", "first_arm", "draft_len", "n_internal", "n_pre",
                                            "n_suf", "n_empty", "wtype")] == [0, 8, 16, 20, 24, 28, 32, 36]
    assert [ARM_DTYPE.fields[n][1] for n in ("off", "len", "reserved")] == [0, 8, 12]


def test_packing_matches_packedseq_layout():
    # reference src/PackedSeq.cpp:45 — PackedSeq<2>: base i in bits 6-2*(i&3) of byte i>>2
    assert pack2("ACGT").tolist() == [0b00011011]
    assert pack2("TGCAT").tolist() == [0b11100100, 0b11000000]
    # reference src/PackedSeq.cpp:48 — PackedSeq<4>: even base in the high nibble, N = 4
    assert pack4("ANT").tolist() == [0x04, 0x30]
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 4, 5, 31, 120):
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))
        assert unpack2(pack2(s), 0, n) == s
        assert unpack4(pack4(s), 0, n) == s
    with pytest.raises(ValueError):
        pack2("ACGN")


def test_batch_select_reindexes_arms():
    specs = [WindowSpec("ACGT", ["AC", "ACG"], ["A"], [], 0, 0), WindowSpec("TTTT", ["TT", "TTT", "T"], [], ["TTTT"], 1, 0),
             WindowSpec("GG", [], [], [], 0, 0)]
    b = build_batch(specs)
    sub = b.select(np.array([1, 2]))
    assert sub.n_win == 2 and sub.n_arms == 4
    assert sub.spec(0).internal == ["TT", "TTT", "T"] and sub.spec(0).suf == ["TTTT"] and sub.spec(0).n_empty == 1
    assert sub.spec(1).draft == "GG"


def test_out_bound_is_host_arithmetic_and_covers_every_consensus():
    """hypo_gpu_out_bound needs no device: it equals the Python mirror of the rule and is an upper bound of
    what the path can emit for every window (checked with the oracle's consensus lengths)."""
    from hypo_b200.synth import edge_case_windows, random_batch
    from tests.oracle_util import oracle_consensus
    lib = ctypes.CDLL(native.LIB_PATH)
    lib.hypo_gpu_out_bound.restype = ctypes.c_uint64
    lib.hypo_gpu_out_bound.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64]
    for b in (build_batch(edge_case_windows()), random_batch(3, 40, kind="mixed", length=60, n_arms=9, err=0.1),
              random_batch(4, 10, kind="internal", length=200, n_arms=6, wtype=1)):
        per = b.out_bound()
        assert lib.hypo_gpu_out_bound(b.win.ctypes.data, b.n_win, b.arms.ctypes.data, b.n_arms) == int(per.sum())
        cons, _ = oracle_consensus(b)
        assert all(len(c) <= int(p) for c, p in zip(cons, per))


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product path must raise, not silently compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(native.HypoGpuError):
        native.init((5, -4, -8, 3, -5, -4), 0)
    b = build_batch([WindowSpec("ACGT", ["ACGT", "ACGT"], [], [], 0, 0)])
    with pytest.raises(native.HypoGpuError):
        native.consensus(b)


def test_options_are_validated_without_a_device():
    """hypo_gpu_set_option is host state: every documented knob accepts its documented range and rejects the rest
    (the execution shapes - group tiers, list ordering, teams, the estimate-driven tier - only change where windows
    run, never a result byte, so they are plain integers)."""
    ok = {"first_tier": (0, 10), "group_tiers": (0, 2), "group_sort": (0, 1), "teams": (0, 1), "big_tier": (0, 1),
          "probe": (0, 1), "scap": (0, 65534)}
    defaults = {"first_tier": 0, "group_tiers": 1, "group_sort": 1, "teams": 1, "big_tier": 1, "probe": 1, "scap": 0}
    try:
        for name, (lo, hi) in ok.items():
            native.set_option(name, lo)
            native.set_option(name, hi)
            for bad in (lo - 1, hi + 1):
                with pytest.raises(native.HypoGpuError):
                    native.set_option(name, bad)
        with pytest.raises(native.HypoGpuError):
            native.set_option("no_such_option", 1)
    finally:
        for name, v in defaults.items():
            native.set_option(name, v)


def test_tier_histogram_call_checks_its_arguments():
    lib = native.lib()
    buf = (ctypes.c_uint32 * 16)(*([7] * 16))
    assert lib.hypo_gpu_last_tier_windows(buf, 16) == 0
    assert list(buf) == [0] * 16          # nothing has run; entries beyond the last tier are zero as documented
    assert lib.hypo_gpu_last_tier_windows(None, 4) != 0
    assert native.N_TIERS == 11 and (native.TIER_QUAD, native.TIER_HALF, native.TIER_BIG) == (8, 9, 10)
