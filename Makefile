# Top-level build.
#   make lib      -> primme_b200/libprimme_b200.so   (product: host C + sm_100a kernels, no CPU path)
#   make oracle   -> oracle/_build/*.so (+ oracle/_ref/libprimme_ref.so when /root/reference exists)
#   make all      -> both
# LAPACK/BLAS for the small host-side dense algebra: any LP64 Fortran-ABI library; the only one in
# this image is the OpenBLAS bundled with opencv.
OB     ?= /opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs
OBLIB  ?= libopenblasp-r0-59ffcd50.3.15.so
LAPACK ?= -L$(OB) -l:$(OBLIB) -Wl,--disable-new-dtags,-rpath,$(OB)
NVCC   ?= nvcc
CC     ?= gcc
ARCH   := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -Xcompiler -fPIC -std=c++17
CFLAGS := -O2 -fPIC -std=c99 -Wall -Wno-unused-function -D_GNU_SOURCE -Werror=incompatible-pointer-types

HOST_SRC := $(wildcard primme_b200/src/*.c)
# the typed sources are compiled twice: plain (dprimme) and with -DPB_COMPLEX (zprimme), like the
# reference's self-including templates (reference src/include/template_types.h:51-204)
TYPED    := davidson dav_ortho dav_project dav_restart dav_jdqmr dav_dynamic dav_refined front hostla
HOST_OBJ := $(patsubst primme_b200/src/%.c,build/host/%.o,$(HOST_SRC)) $(patsubst %,build/host/%_z.o,$(TYPED))
CU_SRC   := $(wildcard primme_b200/csrc/*.cu)
CU_OBJ   := $(patsubst primme_b200/csrc/%.cu,build/cu/%.o,$(CU_SRC))

all: lib oracle

lib: primme_b200/libprimme_b200.so

build/host/%.o: primme_b200/src/%.c primme_b200/src/pb_host.h primme_b200/src/hostla.h include/primme_b200.h include/primme_eigs.h
	@mkdir -p build/host
	$(CC) $(CFLAGS) -c $< -o $@

build/host/%_z.o: primme_b200/src/%.c primme_b200/src/pb_host.h primme_b200/src/hostla.h include/primme_b200.h include/primme_eigs.h
	@mkdir -p build/host
	$(CC) $(CFLAGS) -DPB_COMPLEX -c $< -o $@

build/cu/%.o: primme_b200/csrc/%.cu primme_b200/csrc/pb200_internal.cuh include/primme_b200.h
	@mkdir -p build/cu
	$(NVCC) $(NVFLAGS) -c $< -o $@

primme_b200/libprimme_b200.so: $(HOST_OBJ) $(CU_OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(HOST_OBJ) $(CU_OBJ) -Xlinker --disable-new-dtags -Xlinker -rpath -Xlinker $(OB) -L$(OB) -l:$(OBLIB) -lcudart -ldl -lm

oracle: oracle/_build/libprimme_hostcheck.so
	$(MAKE) -C oracle all
	$(MAKE) examples

# The reference's own examples, compiled UNCHANGED from where they lie under $(REF) against our
# headers and linked against the product library (only when the reference is mounted; the
# binaries are git-ignored under oracle/_ref/ and travel to the GPU box).
REF ?= /root/reference
EXAMPLES := ex_eigs_dseq
examples: $(if $(wildcard $(REF)/examples/ex_eigs_dseq.c),$(addprefix oracle/_ref/examples/,$(EXAMPLES)),)

oracle/_ref/examples/%: $(REF)/examples/%.c primme_b200/libprimme_b200.so
	@mkdir -p oracle/_ref/examples
	$(CC) -O1 -Iinclude $< -o $@ -Lprimme_b200 -lprimme_b200 -Wl,--disable-new-dtags,-rpath,'$$ORIGIN/../../../primme_b200' -Wl,-rpath-link,$(OB) -lm

# host control code linked against the CPU restatement of the kernels: TEST ONLY
oracle/_build/libprimme_hostcheck.so: $(HOST_OBJ) oracle/kernels_ref.c oracle/kernels_ref_z.c oracle/kernels_ref.h
	@mkdir -p oracle/_build
	$(CC) -O2 -fPIC -shared -o $@ $(HOST_OBJ) oracle/kernels_ref.c oracle/kernels_ref_z.c $(LAPACK) -lm

clean:
	rm -rf build primme_b200/libprimme_b200.so oracle/_build

.PHONY: all lib oracle clean examples
