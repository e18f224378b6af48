#!/bin/bash
# TEST INFRASTRUCTURE ONLY (developer tool).  Runs the reference CLI (built by build_reference_cli.sh) on a
# dataset made by make_dataset.py and collects what the reference itself says about its windows:
#   <dataset>/aux/inspect_<contig>.txt   per-window dump written by the reference's own
#                                        Contig::generate_inspect_file (reference src/Contig.cpp:368-453)
#   <dataset>/polished.fa / polished_dump.fa   output of the unmodified and of the dump-enabled binary
#                                        (must be byte-identical)
#   <dataset>/hypo.log                   the reference's Monitor lines ("POA of windows" time among them)
# usage: capture.sh <dataset dir> [threads] [cli dir]
set -euo pipefail
D=$(realpath "$1"); T=${2:-$(nproc)}; CLI=${3:-/tmp/hypo_cli}
ARGS=$(cat "$D/cli_args.txt")
LR=""; [ -f "$D/lr.sam" ] && LR="-B lr.sam"
cd "$D"
[ -f reads.fq ] || printf "@r\nA\n+\nI\n" > reads.fq     # -r must exist; it is only read by the skipped KMC stage
rm -f aux/inspect_*.txt aux/regions.bed
"$CLI/hypo_dump" -r reads.fq -d draft.fa -b sr.sam $LR $ARGS -t "$T" -i -o polished_dump.fa > hypo_dump.log 2>&1
printf 'a b c 1\n' > aux/stage.txt     # the run appended its own stage lines; start behind KMC again
"$CLI/hypo" -r reads.fq -d draft.fa -b sr.sam $LR $ARGS -t "$T" -i -o polished.fa > hypo.log 2>&1
printf 'a b c 1\n' > aux/stage.txt
cmp polished.fa polished_dump.fa && echo "polished output identical with and without the dump"
grep -h "POA\|Number of" hypo.log | head -20
ls -la aux/inspect_*.txt | head
