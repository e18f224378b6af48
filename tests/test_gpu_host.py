"""GPU tests of the host-side drop-in (hypo::Window mirror + WindowBatch packer) and of the
device-resident entry points used by bench.py."""
import numpy as np
import pytest

from hypo_b200 import native
from hypo_b200.batch import build_batch
from hypo_b200.hostlib import host_run, synth_batch
from hypo_b200.synth import edge_case_windows
from tests.oracle_util import DEFAULT_SCORES, oracle_consensus

pytestmark = pytest.mark.gpu


def test_window_mirror_public_api_matches_oracle():
    """Windows built through Window::add_* and polished by Window::generate_consensus_batch
    (the replacement of reference src/Hypo.cpp:236-248) give the oracle's consensus."""
    for b in (build_batch(edge_case_windows()), synth_batch(3, 64, 80, 20, "mixed"),
              synth_batch(4, 16, 250, 12, "internal", wtype=1)):
        want, _ = oracle_consensus(b)
        assert host_run(b, DEFAULT_SCORES, 0) == want


def test_device_resident_path_and_compaction():
    import torch
    native.init(DEFAULT_SCORES, 0)
    b = synth_batch(9, 3000, 120, 30, "internal")
    want, _ = oracle_consensus(b)
    dev = torch.device("cuda", 0)

    def to_dev(x):
        return torch.from_numpy(x.view(np.uint8).reshape(-1)).to(dev)

    bound = b.out_bound()
    pos = np.concatenate([[0], np.cumsum(bound)[:-1]]).astype(np.uint64)
    d_win, d_arms, d_packed, d_pos = to_dev(b.win), to_dev(b.arms), to_dev(b.packed), to_dev(pos)
    d_scr = torch.empty(int(bound.sum()) + 16, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(b.n_win, dtype=torch.int32, device=dev)
    cap = int(bound.sum())
    d_cmp = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(b.n_win + 1, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    native.consensus_batch_device(d_win.data_ptr(), b.n_win, d_arms.data_ptr(), b.n_arms, d_packed.data_ptr(),
                                  b.packed.size, d_scr.data_ptr(), d_pos.data_ptr(), d_len.data_ptr(), s)
    total = native.compact_device(d_scr.data_ptr(), d_pos.data_ptr(), d_len.data_ptr(), b.n_win, d_cmp.data_ptr(),
                                  cap, d_off.data_ptr(), s)
    torch.cuda.synchronize()
    off = d_off.cpu().numpy()
    raw = d_cmp[:total].cpu().numpy().tobytes()
    got = [raw[int(off[i]):int(off[i + 1])].decode() for i in range(b.n_win)]
    assert got == want
    ms, n, tiers = native.last_timing()
    assert ms > 0 and n >= 1 and sum(tiers) >= b.n_win


def test_headline_shape_full_size_properties():
    """BASELINE config-2 shape at a size the oracle cannot finish quickly: check size-independent
    properties (every consensus is ACGT-only, length within the indel envelope of the truth, and
    identical arms => identical consensus) plus an oracle check on a random subsample."""
    native.init(DEFAULT_SCORES, 0)
    b = synth_batch(11, 60000, 120, 30, "internal")
    out, off = native.consensus_batch_host(b)
    lens = np.diff(off.astype(np.int64))
    assert lens.min() >= 100 and lens.max() <= 140
    assert set(np.unique(out[: int(off[-1])]).tolist()) <= {65, 67, 71, 84}
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(b.n_win, 300, replace=False))
    want, _ = oracle_consensus(b.select(idx))
    raw = out.tobytes()
    got = [raw[int(off[i]):int(off[i + 1])].decode() for i in idx]
    assert got == want


def test_pipelined_host_copy_and_its_fallback():
    """>= 65536 windows: the host-buffer entry point copies a head of the batch first and runs it while
    the tail is still being copied.  Same bytes as the oracle; a batch that is NOT laid out in window
    order (permuted windows over a shared slab) silently takes the single-copy path - same result."""
    import numpy as np
    from hypo_b200 import native
    from hypo_b200.hostlib import synth_batch
    from tests.oracle_util import DEFAULT_SCORES, oracle_consensus
    native.init(DEFAULT_SCORES, 0)
    b = synth_batch(77, 70000, 14, 9, "mixed", 0.03)
    want, _ = oracle_consensus(b)
    assert native.consensus(b) == want
    perm = np.random.default_rng(5).permutation(b.n_win)
    got = native.consensus(b.select(perm))
    assert got == [want[i] for i in perm]
