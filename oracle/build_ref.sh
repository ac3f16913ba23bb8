#!/usr/bin/env bash
# Oracle build recipe (test infrastructure only).
# Compiles the reference's own `index`, `overlap`, `rmdup`, `correct` and `assemble` translation units IN PLACE from $SIGA_REF (default /root/reference)
# against the header shim in oracle/shim/ (Boost and log4cxx are not installed here) into oracle/_ref/siga.
# No reference source is copied into this repo; only the binary lands in oracle/_ref/ (git-ignored, travels via gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SIGA_REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: reference sources not found at $REF (ok on the GPU box: prebuilt oracle/_ref is used)" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
CXX="${ORACLE_CXX:-/usr/bin/g++}"  # the image exports CXX=/opt/gcc/bin/g++, which lacks libgomp.spec
FLAGS="-std=c++11 -O2 -fopenmp -w -DHAVE_CONFIG_H -include cstdint -include cassert -include functional -include algorithm -include numeric -I$HERE/shim -I$REF/src -I$REF/3rdparty"
TUS="main indexer overlap rmdup correct correct_processor match assembler bigraph bigraph_visitors bigraph_search suffix_array_builder suffix_array bwt fmindex overlap_builder asqg coord kseq reads"
pids=()
for tu in $TUS; do
  if [ ! -f "$OUT/obj/$tu.o" ] || [ "$REF/src/$tu.cpp" -nt "$OUT/obj/$tu.o" ]; then
    $CXX $FLAGS -c "$REF/src/$tu.cpp" -o "$OUT/obj/$tu.o" &
    pids+=($!)
  fi
done
if [ ! -f "$OUT/obj/utils_zlib.o" ] || [ "$HERE/shim/utils_zlib.cpp" -nt "$OUT/obj/utils_zlib.o" ]; then
  $CXX $FLAGS -c "$HERE/shim/utils_zlib.cpp" -o "$OUT/obj/utils_zlib.o" &
  pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=""
for tu in $TUS utils_zlib; do OBJS="$OBJS $OUT/obj/$tu.o"; done
$CXX -fopenmp -o "$OUT/siga" $OBJS -lz -lpthread
echo "built $OUT/siga"
