#!/bin/bash
# Builds oracle/_ref/ from the reference sources WHERE THEY LIE (/root/reference, read-only); nothing is
# copied into the repository.  Outputs (git-ignored, but shipped to the GPU box with the snapshot):
#   oracle/_ref/librntimgr_ref.so  -- the reference's RNTIManager/Histogram/Interval with its C wrappers
#   oracle/_ref/libfalcon_walk.so  -- oracle/falcon_walk.cc (our restatement of DCISearch.cc) linked with
#                                     those same reference objects and the C oracle
# The rest of the reference's hot path cannot be built: it needs srsRAN (fetched by cmake at configure
# time, absent here) -- see DESIGN.md.
set -e
cd "$(dirname "$0")/.."
REF=/root/reference
[ -d "$REF" ] || { echo "no $REF: keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p oracle/_ref
SRC="$REF/lib/src/util/RNTIManager.cc $REF/lib/src/util/Histogram.cc $REF/lib/src/util/Interval.cc"
INC="-I$REF/lib/include -I$REF/lib/include/falcon/util -Ioracle/ref_stub"
CXXFLAGS="-O2 -g -fPIC -std=c++11 -w"
if [ ! -f oracle/_ref/librntimgr_ref.so ] || [ oracle/build_ref.sh -nt oracle/_ref/librntimgr_ref.so ]; then
  g++ $CXXFLAGS -shared -o oracle/_ref/librntimgr_ref.so $SRC $INC
fi
make -s oracle/liblteoracle.so
if [ ! -f oracle/_ref/libfalcon_walk.so ] || [ oracle/falcon_walk.cc -nt oracle/_ref/libfalcon_walk.so ] || [ oracle/liblteoracle.so -nt oracle/_ref/libfalcon_walk.so ]; then
  g++ -O2 -g -fPIC -std=c++14 -w -ffp-contract=off -shared -o oracle/_ref/libfalcon_walk.so oracle/falcon_walk.cc $SRC $INC -Ioracle \
      -Loracle -llteoracle -Wl,-rpath,'$ORIGIN/..'
fi
echo "oracle/_ref built"
