"""ctypes binding of hypo_b200/libhypo_host.so: the C++ host side of the drop-in (Window /
PackedSeq mirror + batch packer) and the fast seeded synthetic window generator."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence

import numpy as np

from .batch import ARM_DTYPE, WIN_DTYPE, WindowBatch, split_consensus
from .native import HypoGpuError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhypo_host.so")
KINDS = {"internal": 0, "backbone": 1, "prefix": 2, "suffix": 3, "mixed": 4}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HypoGpuError(-1, f"{LIB_PATH} is missing - run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        L.hypo_synth_packed_bound.restype = C.c_uint64
        L.hypo_synth_packed_bound.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.hypo_synth_generate.restype = C.c_int
        L.hypo_synth_generate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_double,
                                          C.c_double, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.POINTER(C.c_uint64), C.c_int]
        L.hypo_host_run.restype = C.c_int
        L.hypo_host_run.argtypes = [C.POINTER(C.c_int8), C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def synth_batch(seed: int, n_win: int, length: int = 120, n_arms: int = 30, kind: str = "internal",
                err: float = 0.01, draft_err: float = 0.03, wtype: int = 0, threads: int = 0) -> WindowBatch:
    """Seeded synthetic windows: truth iid ACGT; draft = truth with draft_err
    sub/ins/del each; arms = (slices of) truth with err sub/ins/del each."""
    L = lib()
    cap = int(L.hypo_synth_packed_bound(n_win, length, n_arms))
    win = np.zeros(n_win, WIN_DTYPE)
    arms = np.zeros(n_win * n_arms, ARM_DTYPE)
    packed = np.zeros(cap, np.uint8)
    used = C.c_uint64(0)
    rc = L.hypo_synth_generate(seed, n_win, length, n_arms, KINDS[kind], err, draft_err, wtype, win.ctypes.data,
                               arms.ctypes.data, packed.ctypes.data, cap, C.byref(used), threads)
    if rc != 0:
        raise RuntimeError("hypo_synth_generate: packed buffer too small")
    packed = packed[: int(used.value) + 16].copy()   # 16 bytes of slack, zero-filled
    return WindowBatch(win, arms, packed,
                       {"seed": seed, "length": length, "n_arms": n_arms, "kind": kind, "err": err,
                        "draft_err": draft_err, "wtype": wtype})


def host_run(batch: WindowBatch, scores: Sequence[int] = (5, -4, -8, 3, -5, -4), device: int = 0) -> List[str]:
    """Builds hypo::Window objects through the mirror's public API and runs
    Window::generate_consensus_batch (the drop-in for reference src/Hypo.cpp:236-248)."""
    L = lib()
    cap = int(batch.out_bound().sum()) + 16
    out = np.empty(cap, np.uint8)
    off = np.zeros(batch.n_win + 1, np.uint64)
    sc = (C.c_int8 * 6)(*[int(x) for x in scores])
    rc = L.hypo_host_run(sc, device, batch.win.ctypes.data, batch.n_win, batch.arms.ctypes.data,
                         batch.packed.ctypes.data, out.ctypes.data, cap, off.ctypes.data)
    if rc != 0:
        raise HypoGpuError(rc, "hypo_host_run failed")
    return split_consensus(out, off)
