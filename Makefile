# Top-level build of the TEST infrastructure (sim + oracle).  The product (CUDA) is built by
# ltesniffer_b200/build.py / __graft_entry__.build().
CC ?= gcc
CXX ?= g++
CFLAGS = -O3 -g -fPIC -std=gnu11 -Wall -Wno-unused-function -ffp-contract=off -fno-fast-math
CXXFLAGS = -O2 -g -fPIC -std=c++17 -Wall -ffp-contract=off -fno-fast-math

all: sim/libltesim.so oracle/liblteoracle.so

sim/libltesim.so: sim/lte_common.c sim/lte_sim.c sim/lte_common.h sim/lte_sim.h include/lte_tables.h
	$(CC) $(CFLAGS) -shared -o $@ sim/lte_common.c sim/lte_sim.c -lm

oracle/liblteoracle.so: oracle/lte_oracle.c oracle/lte_oracle.h sim/lte_common.c sim/lte_common.h include/lte_tables.h
	$(CC) $(CFLAGS) -shared -o $@ oracle/lte_oracle.c sim/lte_common.c -lm -lpthread

clean:
	rm -f sim/*.so oracle/*.so
