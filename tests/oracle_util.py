"""ctypes access to the CHECKERS: oracle/libpoa_oracle.so (C restatement) and, when it was
built in the authoring container, oracle/_ref/libhypo_ref.so (the compiled, unmodified
reference).  Test infrastructure only — nothing under hypo_b200/ imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

from hypo_b200.batch import WindowBatch, WindowSpec, build_batch, split_consensus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
DEFAULT_SCORES = (5, -4, -8, 3, -5, -4)  # reference src/main.cpp defaults, SURVEY.md §8

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def build_oracle() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libpoa_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        _oracle = C.CDLL(path)
        _oracle.poa_oracle_consensus_batch.restype = C.c_int
        _oracle.poa_oracle_spoa_consensus.restype = C.c_int
        _oracle.poa_oracle_window_stats.restype = C.c_int
    return _oracle


_ref = {}


def ref_lib(simd: bool = False):
    """The compiled reference, or None when oracle/_ref was not built (no /root/reference)."""
    if simd not in _ref:
        path = os.path.join(ORACLE_DIR, "_ref", "libhypo_ref_simd.so" if simd else "libhypo_ref.so")
        if not os.path.exists(path):
            _ref[simd] = None
        else:
            lib = C.CDLL(path)
            lib.hypo_ref_consensus_batch.restype = C.c_int
            lib.hypo_ref_spoa_consensus.restype = C.c_int
            assert lib.hypo_ref_is_simd() == int(simd)
            _ref[simd] = lib
    return _ref[simd]


def _scores(scores) -> np.ndarray:
    return np.asarray(scores, dtype=np.int8)


def oracle_consensus(batch: WindowBatch, scores=DEFAULT_SCORES, threads: int = 0) -> Tuple[List[str], float]:
    lib = oracle_lib()
    cap = int(batch.out_bound().sum()) + 16
    out = np.zeros(cap, np.uint8)
    off = np.zeros(batch.n_win + 1, np.uint64)
    sec = C.c_double(0)
    sc = _scores(scores)
    rc = lib.poa_oracle_consensus_batch(
        _ptr(sc, C.POINTER(C.c_int8)), _ptr(batch.win, C.c_void_p), C.c_uint64(batch.n_win),
        _ptr(batch.arms, C.c_void_p), C.c_uint64(batch.n_arms), _ptr(batch.packed, _u8p),
        C.c_uint64(batch.packed.size), _ptr(out, C.c_char_p), C.c_uint64(cap), _ptr(off, _u64p),
        C.c_int(threads), C.byref(sec))
    if rc != 0:
        raise RuntimeError(f"poa_oracle_consensus_batch failed rc={rc}")
    return split_consensus(out, off), sec.value


def ref_consensus(batch: WindowBatch, scores=DEFAULT_SCORES, simd: bool = False, threads: int = 0,
                  schedule: int = 1) -> Tuple[List[str], np.ndarray, float]:
    lib = ref_lib(simd)
    assert lib is not None, "oracle/_ref not built"
    cap = int(batch.out_bound().sum()) + 16
    out = np.zeros(cap, np.uint8)
    off = np.zeros(batch.n_win + 1, np.uint64)
    acc = np.zeros(max(batch.n_arms, 1), np.uint8)
    sec = C.c_double(0)
    sc = _scores(scores)
    rc = lib.hypo_ref_consensus_batch(
        _ptr(sc, C.POINTER(C.c_int8)), _ptr(batch.win, C.c_void_p), C.c_uint64(batch.n_win),
        _ptr(batch.arms, C.c_void_p), C.c_uint64(batch.n_arms), _ptr(batch.packed, _u8p),
        C.c_uint64(batch.packed.size), _ptr(out, C.c_char_p), C.c_uint64(cap), _ptr(off, _u64p),
        _ptr(acc, _u8p), C.c_int(threads), C.c_int(schedule), C.byref(sec))
    if rc != 0:
        raise RuntimeError(f"hypo_ref_consensus_batch failed rc={rc}")
    return split_consensus(out, off), acc[: batch.n_arms].astype(bool), sec.value


def drop_rejected_arms(batch: WindowBatch, accepted: np.ndarray) -> WindowBatch:
    """LONG windows run hypo::Filter::is_good on every arm at insert time (reference
    include/Window.hpp:66-101) — upstream of the hot path.  Rebuild the batch with only the
    arms the reference's Window actually kept, so all three implementations see the same
    Window contents."""
    specs = []
    for w in range(batch.n_win):
        s = batch.spec(w)
        a0 = int(batch.win[w]["first_arm"])
        ni, npre = len(s.internal), len(s.pre)
        specs.append(WindowSpec(
            s.draft,
            [x for i, x in enumerate(s.internal) if accepted[a0 + i]],
            [x for i, x in enumerate(s.pre) if accepted[a0 + ni + i]],
            [x for i, x in enumerate(s.suf) if accepted[a0 + ni + npre + i]],
            s.n_empty, s.wtype))
    return build_batch(specs)


def _spoa(fn, seqs: List[str], m: int, n: int, g: int, types: Optional[List[int]]) -> str:
    blob = "".join(seqs).encode()
    off = np.zeros(len(seqs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    cap = len(blob) + 16
    out = np.zeros(cap, np.uint8)
    out_len = C.c_uint64(0)
    t = np.asarray(types, np.uint8) if types is not None else None
    rc = fn(C.c_int8(m), C.c_int8(n), C.c_int8(g), C.c_char_p(blob), _ptr(off, _u64p),
            C.c_uint32(len(seqs)), _ptr(t, _u8p) if t is not None else None,
            _ptr(out, C.c_char_p), C.c_uint64(cap), C.byref(out_len))
    assert rc == 0
    return out[: out_len.value].tobytes().decode()


def oracle_spoa(seqs, m=5, n=-4, g=-8, types=None) -> str:
    return _spoa(oracle_lib().poa_oracle_spoa_consensus, seqs, m, n, g, types)


def ref_spoa(seqs, m=5, n=-4, g=-8, types=None) -> str:
    return _spoa(ref_lib().hypo_ref_spoa_consensus, seqs, m, n, g, types)


def oracle_stats(batch: WindowBatch, w: int, scores=DEFAULT_SCORES) -> np.ndarray:
    st = np.zeros(5, np.uint64)
    sc = _scores(scores)
    one = batch.win[w : w + 1]
    oracle_lib().poa_oracle_window_stats(
        _ptr(sc, C.POINTER(C.c_int8)), _ptr(one, C.c_void_p), _ptr(batch.arms, C.c_void_p),
        _ptr(batch.packed, _u8p), _ptr(st, _u64p))
    return st


def oracle_growth(batch: WindowBatch, w: int, scores=DEFAULT_SCORES) -> np.ndarray:
    """[n_sequences, 3] nodes, edges, cumulative DP cells after every sequence of window w (both rounds of a
    LONG window, one after the other)."""
    one = batch.win[w : w + 1]
    cap = 2 * (int(one["n_internal"][0]) + int(one["n_pre"][0]) + int(one["n_suf"][0]) + 2)
    g = np.zeros(3 * cap, np.uint64)
    n = C.c_uint32(0)
    sc = _scores(scores)
    oracle_lib().poa_oracle_window_growth(
        _ptr(sc, C.POINTER(C.c_int8)), _ptr(one, C.c_void_p), _ptr(batch.arms, C.c_void_p),
        _ptr(batch.packed, _u8p), _ptr(g, _u64p), C.c_uint32(cap), C.byref(n))
    return g[: 3 * n.value].reshape(-1, 3).astype(np.int64)
