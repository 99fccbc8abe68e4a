# Builds libpistoseg_b200.so (sm_100a only) in-tree.  `python -c "import __graft_entry__ as g; g.build()"` calls this.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
EXTRA     ?=
BUILD     ?= build
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Ipistoseg_b200/csrc -diag-suppress 128 $(EXTRA)
SRCS      := $(wildcard pistoseg_b200/csrc/*.cu)
OBJS      := $(patsubst pistoseg_b200/csrc/%.cu,$(BUILD)/%.o,$(SRCS))
LIB       ?= pistoseg_b200/libpistoseg_b200.so

all: $(LIB)

$(BUILD)/%.o: pistoseg_b200/csrc/%.cu $(wildcard pistoseg_b200/csrc/*.cuh) include/pistoseg_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

clean:
	rm -rf build $(LIB)

.PHONY: all clean
