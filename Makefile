# Builds the C-ABI library (sm_100a only) and the CPU oracle (test infrastructure).
NVCC ?= nvcc
NVFLAGS = -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude
SRCS = $(wildcard pyranges_b200/csrc/*.cu)
OBJS = $(patsubst pyranges_b200/csrc/%.cu,build/%.o,$(SRCS))
LIB = pyranges_b200/libpyranges_b200.so

all: $(LIB) oracle/libiv_oracle.so

build/%.o: pyranges_b200/csrc/%.cu pyranges_b200/csrc/common.cuh pyranges_b200/csrc/query.cuh include/pyranges_b200.h
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS)

oracle/libiv_oracle.so: oracle/iv_oracle.c
	gcc -O3 -march=x86-64-v2 -fPIC -shared -o $@ $<

clean:
	rm -rf build $(LIB) oracle/libiv_oracle.so
.PHONY: all clean
