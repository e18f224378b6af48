#!/usr/bin/env python
"""Replays a window stream captured from the reference CLI (aux/inspect_<contig>.txt, written by
Contig::generate_inspect_file when reference src/Hypo.cpp:262,265,271 are enabled) on the GPU and
compares every window's consensus with the one the reference recorded.

  python tools/replay_inspect.py aux/inspect_ctg1.txt [--out replayed.txt] [--scores 5 -4 -8 3 -5 -4]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hypo_b200.hostlib import InspectStream  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("paths", nargs="+")
    ap.add_argument("--scores", type=int, nargs=6, default=[5, -4, -8, 3, -5, -4])
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out", default="", help="write the stream back with the GPU consensus (single input only)")
    a = ap.parse_args()
    worst = 0
    for p in a.paths:
        s = InspectStream(p)
        bad, sec = s.replay(a.scores, a.device, a.out if len(a.paths) == 1 else "")
        worst = max(worst, bad)
        print(json.dumps({"file": p, "regions": s.n_regions, "windows": s.n_windows, "polished_bp": s.polished_bp,
                          "seconds": sec, "mbp_per_s": s.polished_bp / 1e6 / max(sec, 1e-9),
                          "windows_differing_from_recorded": bad}))
        s.close()
    sys.exit(1 if worst else 0)


if __name__ == "__main__":
    main()
