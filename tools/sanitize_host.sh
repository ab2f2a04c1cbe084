#!/bin/bash
# Runs the CPU test suite against a build of libltephy_b200.so whose HOST objects (lte_host, host_search, sinks, harq) are compiled with
# AddressSanitizer or UBSan (usage: tools/sanitize_host.sh address|undefined).  The CUDA objects are the ones already built in
# ltesniffer_b200/_obj; the regular library is put back afterwards.  Needs no GPU.
set -e
cd "$(dirname "$0")/.."
KIND=${1:-address}
T=$(mktemp -d)
for f in lte_host host_search sinks harq; do
  g++ -O1 -g -std=c++17 -fPIC -fsanitize=$KIND -fno-omit-frame-pointer -ffp-contract=off -c -o $T/$f.o ltesniffer_b200/csrc/$f.cpp
done
LIBS=$([ "$KIND" = address ] && echo "-lasan" || echo "-lubsan")
nvcc -gencode arch=compute_100a,code=sm_100a --shared -o $T/lib.so ltesniffer_b200/_obj/k_*.o ltesniffer_b200/_obj/ltephy_capi.cu.o ltesniffer_b200/_obj/shard.cu.o $T/*.o -ldl -Xlinker $LIBS
cp -p ltesniffer_b200/libltephy_b200.so $T/orig.so
trap 'cp -p $T/orig.so ltesniffer_b200/libltephy_b200.so' EXIT
cp $T/lib.so ltesniffer_b200/libltephy_b200.so && touch ltesniffer_b200/libltephy_b200.so
PRE=$([ "$KIND" = address ] && gcc -print-file-name=libasan.so || echo "")
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 LD_PRELOAD=$PRE \
  python -m pytest tests -q -m "not gpu" -p no:cacheprovider --deselect tests/test_sharding_gloo.py --deselect tests/test_bench_contract.py \
  --deselect tests/test_example_cpp.py --deselect tests/test_compat_shim.py --deselect tests/test_capi_exports.py
