"""Window streams in the reference's dump format (hypo_b200/host/WindowStream.hpp).

The golden stream tests/golden/inspect_ref.txt.gz was printed by the reference's own
Window::operator<< after the reference computed each consensus (tests/golden/make_golden.py)."""
import gzip
import os

import numpy as np
import pytest

from hypo_b200.batch import build_batch
from hypo_b200.hostlib import InspectStream, write_inspect
from hypo_b200.synth import edge_case_windows, random_batch
from tests.oracle_util import DEFAULT_SCORES, oracle_consensus

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def golden_path(tmp_path):
    p = tmp_path / "inspect_ctg_golden.txt"
    p.write_bytes(gzip.open(os.path.join(HERE, "golden", "inspect_ref.txt.gz")).read())
    return str(p)


def test_reference_stream_parses_and_pins_the_oracle(golden_path):
    s = InspectStream(golden_path)
    assert s.n_regions == 154 and s.n_windows == 71
    assert s.polished_bp == int(s.batch.win["draft_len"].sum())
    got, _ = oracle_consensus(s.batch, DEFAULT_SCORES)
    assert got == s.recorded          # the C restatement agrees with what the reference recorded
    assert (s.batch.win["wtype"] == 1).sum() >= 6   # the LNG regions come back as LONG windows


def test_write_read_round_trip(tmp_path):
    b = build_batch(edge_case_windows())
    b2 = random_batch(5, 40, kind="mixed", length=60, n_arms=14, err=0.04)
    for k, batch in enumerate((b, b2)):
        cons, _ = oracle_consensus(batch)
        p = str(tmp_path / f"inspect_{k}.txt")
        write_inspect(p, batch, cons, contig=f"ctg{k}")
        s = InspectStream(p)
        # windows without any arm and without empties are plain regions in the dump
        keep = [w for w in range(batch.n_win)
                if int(batch.win["n_internal"][w] + batch.win["n_pre"][w] + batch.win["n_suf"][w] + batch.win["n_empty"][w]) > 0]
        assert s.n_windows == len(keep) and s.n_regions == 2 * batch.n_win
        assert s.recorded == [cons[w] for w in keep]
        ref = batch.select(np.array(keep))
        for w in range(s.n_windows):
            assert s.batch.spec(w) == ref.spec(w)


def test_malformed_stream_is_rejected(tmp_path):
    p = tmp_path / "bad.txt"
    p.write_text(">c\n#1\n==========(0-3)\tOTH\t1\t0\t0\t0\n++\tACGT\n")
    with pytest.raises(ValueError):
        InspectStream(str(p))
    p.write_text(">c\n#0\n==========(0-3)\tSR\t0\t0\t0\t0\n++\tACGT\n++\tACGT\n")
    with pytest.raises(ValueError):
        InspectStream(str(p))
    # more regions announced than printed is the reference's own format once long reads are used
    # (regions swallowed by a LONG pseudo-window are not printed, reference src/Contig.cpp:424-451)
    p.write_text(">c\n#2\n==========(0-3)\tSR\t0\t0\t0\t0\n++\tACGT\n++\tACGT\n")
    assert InspectStream(str(p)).n_regions == 1
    with pytest.raises(ValueError):
        InspectStream(str(tmp_path / "missing.txt"))


# Streams captured from the reference COMMAND-LINE PROGRAM (tests/golden/make_cli_streams.sh): every window
# the real pipeline made of a seeded 60 kb genome - solid-k-mer segmentation, arms cut out of aligned
# reads, pruning - with the consensus the reference computed for it.
CLI_STREAMS = (("cli_short_60kb.inspect.gz", 3443, 1720, 0), ("cli_long_60kb.inspect.gz", 2871, 1443, 25))


def _cli_path(tmp_path, name):
    p = tmp_path / name.replace(".gz", ".txt")
    p.write_bytes(gzip.open(os.path.join(HERE, "golden", name)).read())
    return str(p)


@pytest.mark.parametrize("name,n_regions,n_windows,n_long", CLI_STREAMS)
def test_cli_captured_streams_pin_the_oracle(tmp_path, name, n_regions, n_windows, n_long):
    s = InspectStream(_cli_path(tmp_path, name))
    assert (s.n_regions, s.n_windows) == (n_regions, n_windows)
    assert int((s.batch.win["wtype"] == 1).sum()) == n_long
    got, _ = oracle_consensus(s.batch, DEFAULT_SCORES)
    assert got == s.recorded


def _polished(name):
    """The sequence line of the reference CLI's own output FASTA for the captured run."""
    txt = gzip.open(os.path.join(HERE, "golden", name.replace(".inspect.gz", ".polished.fa.gz"))).read().decode()
    lines = txt.split("\n")
    assert lines[0] == ">ctg1"
    return lines[1]


@pytest.mark.parametrize("name,n_regions,n_windows,n_long", CLI_STREAMS)
def test_cli_captured_streams_stitch_to_the_reference_output(tmp_path, name, n_regions, n_windows, n_long):
    """Regions + recorded consensus strings, stitched like Contig::operator<< (reference src/Contig.cpp:345-366),
    give exactly the polished FASTA the reference CLI wrote in the same run."""
    s = InspectStream(_cli_path(tmp_path, name))
    assert s.stitched(3) == _polished(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_regions,n_windows,n_long", CLI_STREAMS)
def test_device_stitching_reproduces_the_reference_cli_output(tmp_path, name, n_regions, n_windows, n_long):
    """windows polished on the GPU, contig stitched on the GPU (hypo_gpu_stitch) - from
    consensus bytes sent back in, and straight from the result still resident on the device - equals the
    reference CLI's polished FASTA byte for byte."""
    s = InspectStream(_cli_path(tmp_path, name))
    bad, _ = s.replay(DEFAULT_SCORES, 0)
    assert bad == 0
    want = _polished(name)
    assert s.stitched(0) == want
    assert s.stitched(1) == want
    assert s.stitched(2) == want


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_regions,n_windows,n_long", CLI_STREAMS)
def test_replay_cli_captured_streams_on_the_gpu(tmp_path, name, n_regions, n_windows, n_long):
    path = _cli_path(tmp_path, name)
    s = InspectStream(path)
    out = str(tmp_path / "replayed.txt")
    bad, sec = s.replay(DEFAULT_SCORES, 0, out)
    assert bad == 0 and sec > 0
    assert open(out, "rb").read() == open(path, "rb").read()   # byte-identical re-dump


@pytest.mark.gpu
def test_replay_reference_stream_on_the_gpu(golden_path, tmp_path):
    s = InspectStream(golden_path)
    out = str(tmp_path / "replayed.txt")
    bad, sec = s.replay(DEFAULT_SCORES, 0, out)
    assert bad == 0 and sec > 0
    # written back with the GPU consensus, the stream is byte-identical to the reference's dump
    assert open(out, "rb").read() == open(golden_path, "rb").read()


@pytest.mark.gpu
def test_replay_detects_a_wrong_recorded_consensus(tmp_path):
    batch = random_batch(9, 32, kind="internal", length=50, n_arms=10)
    cons, _ = oracle_consensus(batch)
    cons[3] = cons[3] + "A"
    p = str(tmp_path / "tampered.txt")
    write_inspect(p, batch, cons)
    bad, _ = InspectStream(p).replay(DEFAULT_SCORES, 0)
    assert bad == 1
