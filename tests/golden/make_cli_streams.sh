#!/bin/bash
# Regenerates tests/golden/cli_{short,long}_60kb.inspect.gz: window streams captured from the REFERENCE
# command-line program itself (built from /root/reference by tools/capture/build_reference_cli.sh; the
# per-window dump is the reference's own Contig::generate_inspect_file, src/Contig.cpp:368-453) on
# seeded synthetic genomes (tools/capture/make_dataset.py).  Run in the authoring container only
# (needs /root/reference); the fixtures travel, the reference does not.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd); W=${1:-/tmp/hypo_capture}
"$ROOT/tools/capture/build_reference_cli.sh" /tmp/hypo_cli
python "$ROOT/tools/capture/make_dataset.py" "$W/short" --size 60000 --cov 40 --seed 3
"$ROOT/tools/capture/capture.sh" "$W/short" 8
python "$ROOT/tools/capture/make_dataset.py" "$W/long" --size 60000 --cov 30 --long 25 --draft-err 0.04 --sr-gaps 0.15 --seed 5
"$ROOT/tools/capture/capture.sh" "$W/long" 8
gzip -9 -n -c "$W/short/aux/inspect_ctg1.txt" > "$HERE/cli_short_60kb.inspect.gz"
gzip -9 -n -c "$W/long/aux/inspect_ctg1.txt" > "$HERE/cli_long_60kb.inspect.gz"
gzip -9 -n -c "$W/short/polished.fa" > "$HERE/cli_short_60kb.polished.fa.gz"   # the reference CLI's own output
gzip -9 -n -c "$W/long/polished.fa" > "$HERE/cli_long_60kb.polished.fa.gz"
gzip -9 -n -c "$W/short/sr.sam" > "$HERE/cli_short_60kb.sam.gz"   # the alignments the run read (input of arm extraction)
# support counters of the same run (tools/capture: hypo_dump2 = the dump-enabled CLI + two more dumps)
(cd "$W/short" && printf 'a b c 1\n' > aux/stage.txt && /tmp/hypo_cli/hypo_dump2 -r reads.fq -d draft.fa -b sr.sam $(cat cli_args.txt) -t 8 -i -o polished_dump2.fa > hypo_dump2.log 2>&1 && cmp polished.fa polished_dump2.fa)
gzip -9 -n -c "$W/short/aux/kmer_support_ctg1.txt" > "$HERE/cli_short_60kb.kmer_support.gz"
gzip -9 -n -c "$W/short/aux/minimiser_support_ctg1.txt" > "$HERE/cli_short_60kb.minimiser_support.gz"
