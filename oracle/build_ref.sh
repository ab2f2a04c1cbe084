#!/bin/bash
# Builds oracle/_ref/ from the reference sources WHERE THEY LIE (/root/reference, read-only); nothing is
# copied into the repository.  Outputs (git-ignored, but shipped to the GPU box with the snapshot):
#   oracle/_ref/librntimgr_ref.so  -- the reference's RNTIManager/Histogram/Interval with its C wrappers
#   oracle/_ref/libfalcon_walk.so  -- oracle/falcon_walk.cc (our restatement of DCISearch.cc) linked with
#                                     those same reference objects and the C oracle
#   oracle/_ref/libfalcon_ref.so   -- the reference's OWN blind search and DCI -> grant code, unmodified: src/src/{DCISearch,
#                                     DCICollection,MetaFormats,SubframeInfo,SubframePower,SubframeInfoConsumer,ULSchedule,HARQ,
#                                     MCSTracking,PhyCommon,Sniffer_dependency,DCIPrint}.cc, lib/src/phy/falcon_phch/*.c,
#                                     lib/src/phy/falcon_ue/falcon_ue_dl.c, lib/src/util/*.cc, lib/src/prof/*.cc -- compiled against
#                                     the srsRAN-compatible header tree compat/srsran and linked to libltephy_srsran_compat.so,
#                                     behind the C wrapper oracle/ref_walk.cc
# The srsRAN signal processing itself cannot be built: it is fetched by cmake at configure time and absent here.
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
# ---- the reference's own search / grant code against compat/srsran ------------------------------------------------------------
OUT=oracle/_ref/libfalcon_ref.so
COMPAT_LIB=ltesniffer_b200/libltephy_srsran_compat.so
if [ -f "$COMPAT_LIB" ]; then
  NEWEST=$(ls -t oracle/ref_walk.cc oracle/build_ref.sh $COMPAT_LIB $(find compat -name '*.h') | head -1)
  if [ ! -f $OUT ] || [ "$NEWEST" -nt $OUT ]; then
    OBJ=oracle/_ref/obj
    mkdir -p $OBJ
    RINC="-Icompat -Iinclude -I$REF/lib/include -I$REF/lib/include/falcon/util -I$REF/src -I$REF/src/include"
    for f in falcon_pdcch falcon_dci dl_sniffer_pdsch ul_sniffer_pusch; do
      gcc -std=gnu11 -O2 -fPIC -w -c -o $OBJ/$f.o $REF/lib/src/phy/falcon_phch/$f.c $RINC &
    done
    gcc -std=gnu11 -O2 -fPIC -w -c -o $OBJ/falcon_ue_dl.o $REF/lib/src/phy/falcon_ue/falcon_ue_dl.c $RINC &
    for f in RNTIManager Histogram Interval; do g++ -std=c++17 -O2 -fPIC -w -c -o $OBJ/$f.o $REF/lib/src/util/$f.cc $RINC & done
    for f in Lifetime Stopwatch; do g++ -std=c++17 -O2 -fPIC -w -c -o $OBJ/$f.o $REF/lib/src/prof/$f.cc $RINC & done
    for f in DCISearch DCICollection SubframeInfo SubframeInfoConsumer SubframePower ULSchedule HARQ MCSTracking MetaFormats PhyCommon Sniffer_dependency DCIPrint; do
      g++ -std=c++17 -O2 -fPIC -w -c -o $OBJ/$f.o $REF/src/src/$f.cc $RINC &
    done
    g++ -std=c++17 -O2 -fPIC -w -c -o $OBJ/ref_walk.o oracle/ref_walk.cc $RINC &
    wait
    g++ -shared -o $OUT $OBJ/*.o -Lltesniffer_b200 -lltephy_srsran_compat -lltephy_b200 -lpthread -Wl,-rpath,'$ORIGIN/../../ltesniffer_b200' -Wl,--no-undefined
  fi
fi
echo "oracle/_ref built"
