# Builds the C-ABI library for sm_100a and the plain-C consumer; the Python side builds the same
# library through __graft_entry__.build() / rag_arc_b200/_native.py.
NVCC      ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177,550,128
CSRC      := $(sort $(wildcard rag_arc_b200/csrc/*.cu))
LIB       := rag_arc_b200/libragarc_b200.so

.PHONY: lib stats ctest clean
lib: $(LIB)

$(LIB): $(CSRC) rag_arc_b200/csrc/common.cuh include/ragarc_b200.h
	$(NVCC) $(NVCCFLAGS) -shared $(CSRC) -o $@ -ldl

# the same library with the scoring kernel's experiment counters compiled in (benchmarks/tc_stats.py:
# RAGARC_LIB=$(abspath rag_arc_b200/libragarc_b200_stats.so) RAGARC_TC_STATS=1 python benchmarks/tc_stats.py)
stats: $(CSRC) rag_arc_b200/csrc/common.cuh include/ragarc_b200.h
	$(NVCC) $(NVCCFLAGS) -DRAGARC_TC_STATS_BUILD -shared $(CSRC) -o rag_arc_b200/libragarc_b200_stats.so -ldl

# tests/c/index_smoke.c: needs a B200 to pass; without one it exits 2 with RAGARC_ERR_CUDA
ctest: $(LIB)
	$(CC) -std=c99 -Wall -Wextra -pedantic -Werror -I include tests/c/index_smoke.c -o tests/c/index_smoke \
	    -L rag_arc_b200 -lragarc_b200 -Wl,-rpath,$(abspath rag_arc_b200) -lm
	tests/c/index_smoke

clean:
	rm -f $(LIB) rag_arc_b200/libragarc_b200_stats.so tests/c/index_smoke
