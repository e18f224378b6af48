#!/bin/bash
# TEST INFRASTRUCTURE ONLY (developer tool; nothing in tests/, bench.py or smoke() runs it).
#
# Builds the UNMODIFIED reference command-line program, plus a copy whose ONLY change is that the
# three commented-out dump lines of src/Hypo.cpp (262, 265, 271: bedfile + Contig::generate_inspect_file,
# reference src/Contig.cpp:368-453) are enabled, into a scratch directory (default /tmp/hypo_cli).
# Nothing is written to /root/reference and no reference source enters this repository: the sources are
# compiled where they lie, with g++ directly (no cmake, no reference build system).  Vendored htslib is
# copied to the scratch directory because its Makefile writes next to its sources.
#
# Output: $OUT/hypo (unmodified), $OUT/hypo_dump (dump enabled), $OUT/hypo_dump2 (dump + support counters).  SISD spoa engine (no -march), like
# the default build.
set -euo pipefail
REF=${REF:-/root/reference}
OUT=${1:-/tmp/hypo_cli}
J=${J:-$(nproc)}
mkdir -p "$OUT/obj" "$OUT/src"

# --- htslib (static) ---------------------------------------------------------------------------
if [ ! -f "$OUT/htslib/libhts.a" ]; then
    rm -rf "$OUT/htslib"; cp -r "$REF/external/install/htslib" "$OUT/htslib"
    printf '#define HAVE_FSEEKO 1\n#define HAVE_DRAND48 1\n' > "$OUT/htslib/config.h"
    make -C "$OUT/htslib" -j"$J" lib-static CFLAGS="-O2 -w" NONCONFIGURE_OBJS= >/dev/null   # (no libcurl here)
fi

# --- the one generated file of sdsl-lite (a path string for an HTML report, unused here) -------------
sed "s#@PROJECT_SOURCE_DIR@#$REF/external/sdsl-lite#" "$REF/external/sdsl-lite/lib/structure_tree.cpp.cmake" \
    > "$OUT/src/structure_tree.cpp"
# --- Hypo.cpp with the dump lines enabled -------------------------------------------------------------
sed -e '262s#^\( *\)//#\1#' -e '265s#^\( *\)//#\1#' -e '271s#^\( *\)//#\1#' "$REF/src/Hypo.cpp" > "$OUT/src/Hypo_dump.cpp"
if [ "$(diff "$REF/src/Hypo.cpp" "$OUT/src/Hypo_dump.cpp" | grep -c '^>')" != 3 ]; then
    echo "the dump lines of Hypo.cpp are not where SURVEY.md says they are" >&2; exit 1
fi

# --- Hypo.cpp with the dump lines enabled AND two more dumps for the support-counting step:
# after "Solid kmers support update" every solid k-mer's position, id, coverage and support
# (Contig::_kmerinfo, filled by Alignment::update_solidkmers_support, src/Alignment.cpp:65-131), and after
# "Minimisers support update" the region boundaries of that stage and every minimiser's relative position,
# value, coverage and support (Contig::_minimserinfo, src/Alignment.cpp:133-220).  The members are private:
# the access specifiers are lifted for the reference's own headers in this throw-away translation unit only
# (every standard / third-party header is included first).
python3 - "$REF/src/Hypo.cpp" "$OUT/src/Hypo_dump2.cpp" <<'PY'
import sys
src = open(sys.argv[1]).read().split("\n")
pre = ['#include <bits/stdc++.h>', '#include <omp.h>', '#include <zlib.h>', '#include <sdsl/bit_vectors.hpp>', '#include <sdsl/util.hpp>',
       '#include "htslib/sam.h"', '#include "htslib/kseq.h"', '#include "slog/Monitor.hpp"', '#include "suk/SolidKmers.hpp"',
       '#include "spoa/spoa.hpp"', '#define private public']
out = []
for i, l in enumerate(src, 1):
    if i in (262, 265, 271):
        l = l.replace("//", "", 1)
    out.append(l)
    if 'Solid kmers support update. ' in l:
        out.append('        for (UINT64 cid=initial_cid; cid < final_cid; ++cid) { auto& c = *_contigs[cid]; std::ofstream kf(std::string("aux/kmer_support_")+c._name+".txt");'
                   ' for (size_t i=0;i<c._kmerinfo.size();++i) kf << c._Ssolid_pos(i+1) << "\\t" << c._kmerinfo[i]->kid << "\\t" << c._kmerinfo[i]->coverage << "\\t" << c._kmerinfo[i]->support << "\\n"; }')
    if 'Minimisers support update. ' in l:
        out.append('        for (UINT64 cid=initial_cid; cid < final_cid; ++cid) { auto& c = *_contigs[cid]; std::ofstream mf(std::string("aux/minimiser_support_")+c._name+".txt");'
                   ' mf << "even\\t" << (c._is_win_even?1:0) << "\\n"; mf << "bounds";'
                   ' for (size_t p=0;p<c._reg_pos.size();++p) if (c._reg_pos[p]) mf << "\\t" << p; mf << "\\n";'
                   ' for (size_t m=0;m<c._minimserinfo.size();++m) { auto& mi=*c._minimserinfo[m]; mf << "#" << m << "\\t" << mi.rel_pos.size() << "\\n";'
                   ' for (size_t j=0;j<mi.rel_pos.size();++j) mf << mi.rel_pos[j] << "\\t" << mi.minimisers[j] << "\\t" << mi.coverage[j] << "\\t" << mi.support[j] << "\\n"; } }')
open(sys.argv[2], "w").write("\n".join(pre + out))
PY

INC="-I$REF/include -I$REF/src -I$REF/external/spoa/include -I$REF/external/spoa/src -I$OUT/htslib
     -I$REF/external/suk/include -I$REF/external/suk/external/kmc_api -I$REF/external/slog/include
     -I$REF/external/sdsl-lite/include -I$REF/external/slog/src"
CXXFLAGS="-O3 -std=c++17 -fopenmp -w -include cstdint -include stdexcept $INC"

SRCS="$REF/src/Alignment.cpp $REF/src/Contig.cpp $REF/src/main.cpp $REF/src/PackedSeq.cpp $REF/src/Window.cpp
      $REF/external/spoa/src/alignment_engine.cpp $REF/external/spoa/src/graph.cpp
      $REF/external/spoa/src/sisd_alignment_engine.cpp $REF/external/spoa/src/simd_alignment_engine.cpp
      $REF/external/suk/src/SolidKmers.cpp $REF/external/suk/external/kmc_api/kmc_file.cpp
      $REF/external/suk/external/kmc_api/kmer_api.cpp $REF/external/suk/external/kmc_api/mmer.cpp
      $REF/external/slog/src/Monitor.cpp $OUT/src/structure_tree.cpp"
# sdsl-lite: HyPo only uses bit vectors with rank/select support (include/Contig.hpp:27-28); the suffix-array
# construction files need libdivsufsort's cmake-generated header and are not part of what is linked.
for f in bits coder_elias_delta coder_elias_gamma coder_fibonacci config io memory_management ram_filebuf ram_fs \
         rrr_vector_15 sd_vector sfstream uint128_t uint256_t util; do
    SRCS="$SRCS $REF/external/sdsl-lite/lib/$f.cpp"
done

objs=""
pids=()
for f in $SRCS; do
    o="$OUT/obj/$(echo "$f" | md5sum | cut -c1-8)_$(basename "${f%.cpp}").o"
    objs="$objs $o"
    if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then
        g++ $CXXFLAGS -c "$f" -o "$o" &
        pids+=($!)
        if [ "${#pids[@]}" -ge "$J" ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
    fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
g++ $CXXFLAGS -c "$REF/src/Hypo.cpp" -o "$OUT/obj/Hypo.o"
g++ $CXXFLAGS -c "$OUT/src/Hypo_dump.cpp" -o "$OUT/obj/Hypo_dump.o"
g++ $CXXFLAGS -c "$OUT/src/Hypo_dump2.cpp" -o "$OUT/obj/Hypo_dump2.o"
LIBS="$OUT/htslib/libhts.a -lz -lpthread"
g++ -fopenmp -o "$OUT/hypo" "$OUT/obj/Hypo.o" $objs $LIBS
g++ -fopenmp -o "$OUT/hypo_dump" "$OUT/obj/Hypo_dump.o" $objs $LIBS
g++ -fopenmp -o "$OUT/hypo_dump2" "$OUT/obj/Hypo_dump2.o" $objs $LIBS
echo "built $OUT/hypo, $OUT/hypo_dump and $OUT/hypo_dump2"
