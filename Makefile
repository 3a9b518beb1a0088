# megakv_b200 -- build of the B200-native libgpuhash (sm_100a only).
#   make            shared + static library under megakv_b200/lib/
#   make oracle     CPU oracle (test infrastructure)
#   make ref        reference kernels in legacy-warp mode into oracle/_ref/ (needs /root/reference)
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
# MEM_P / policy defaults of the three legacy entry points (run-time geometry via gpuhash_ex.h)
DEFS     ?=
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Iinclude -Imegakv_b200/csrc $(DEFS) \
            -Xcompiler -fPIC -Xcompiler -fno-exceptions -Xcompiler -fno-rtti -Xcompiler -Wall
CSRC     := megakv_b200/csrc
LIBDIR   := megakv_b200/lib
OBJS     := $(LIBDIR)/libgpuhash.o $(LIBDIR)/gpuhash_index.o $(LIBDIR)/gpuhash_workload.o $(LIBDIR)/gpuhash_shard.o $(LIBDIR)/gpuhash_xchg.o $(LIBDIR)/gpuhash_ring.o
HDRS     := include/gpu_hash.h include/libgpuhash.h include/gpuhash_ex.h $(wildcard $(CSRC)/*.cuh)

all: $(LIBDIR)/libgpuhash.so $(LIBDIR)/libgpuhash.a

$(LIBDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

# static cudart: the .so is self-contained whether or not the host process has torch loaded
$(LIBDIR)/libgpuhash.so: $(OBJS)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS)

# same path/name the reference's src/Makefile:26 links with -L$(LIBGPUHASHDIR)lib -lgpuhash
$(LIBDIR)/libgpuhash.a: $(OBJS)
	ar rcs $@ $(OBJS)

# a plain-C host that drives the library like Mega-KV's scheduler (legacy ABI, index_submit, ring); static archive + cudart
example: $(LIBDIR)/libgpuhash.a
	gcc -O2 -std=gnu99 -Wall -Iinclude -I/usr/local/cuda/include examples/scheduler_cycle.c $(LIBDIR)/libgpuhash.a \
	    -L/usr/local/cuda/lib64 -lcudart -lrt -lpthread -ldl -o examples/scheduler_cycle

oracle:
	$(MAKE) -C oracle liboracle.so

ref:
	$(MAKE) -C oracle ref

clean:
	rm -rf $(LIBDIR) ; $(MAKE) -C oracle clean

.PHONY: all example oracle ref clean
