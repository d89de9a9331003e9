# Builds the B200-native HPR-LP engine with the reference's artefact names
# (reference Makefile:114-162): lib/libhprlp.a, lib/libhprlp.so, build/solve_mps_file.
# sm_100a only; -lineinfo so ncu source pages map to kernels.cuh.
CUDA_PATH ?= /usr/local/cuda
NVCC := $(CUDA_PATH)/bin/nvcc
HOSTCXX ?= /usr/bin/g++
SRC := hpr-lp-c_b200/csrc
BUILD := build
LIB := lib
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC \
           -Xcompiler -Wall -Xcompiler -Wno-unused-function -Xcompiler -fopenmp -Iinclude
CXXFLAGS := -O2 -std=c++17 -fPIC -fopenmp -Wall -Iinclude -I$(CUDA_PATH)/include
LIBS := -L$(CUDA_PATH)/lib64 -lcurand -lz -lgomp -lpthread -ldl
RPATH := -Xlinker -rpath -Xlinker $(CUDA_PATH)/lib64

# PSLP presolver (third-party, Apache-2.0): compiled from where its sources lie when present, never copied here
PSLP_DIR ?= /root/reference/third_party/PSLP
ifneq ($(wildcard $(PSLP_DIR)/src/core/Presolver.c),)
PSLP_SRCS := $(filter-out $(PSLP_DIR)/src/core/Debugger.c,$(wildcard $(PSLP_DIR)/src/core/*.c)) $(wildcard $(PSLP_DIR)/src/explorers/*.c)
PSLP_OBJS := $(patsubst $(PSLP_DIR)/src/%.c,$(BUILD)/pslp/%.o,$(PSLP_SRCS))
PSLP_INC := -I$(PSLP_DIR)/include/PSLP -I$(PSLP_DIR)/include/core -I$(PSLP_DIR)/include/data_structures -I$(PSLP_DIR)/include/explorers
CXXFLAGS += -DHPRLP_WITH_PSLP $(PSLP_INC)
else
# prebuilt objects shipped with the snapshot (the GPU box has no reference tree)
PSLP_OBJS := $(wildcard $(BUILD)/pslp/core/*.o) $(wildcard $(BUILD)/pslp/explorers/*.o)
ifneq ($(PSLP_OBJS),)
$(warning PSLP sources not found; reusing prebuilt objects in $(BUILD)/pslp)
endif
endif

CU_SRCS := $(SRC)/engine.cu $(SRC)/api.cu $(SRC)/batched.cu $(SRC)/transpose.cu $(SRC)/partitioned.cu $(SRC)/synth_device.cu $(SRC)/collective.cu
CPP_SRCS := $(SRC)/mps_reader.cpp $(SRC)/presolve.cpp $(SRC)/nccl_shim.cpp $(SRC)/host_utils.cpp
OBJS := $(patsubst $(SRC)/%.cu,$(BUILD)/%.o,$(CU_SRCS)) $(patsubst $(SRC)/%.cpp,$(BUILD)/%.o,$(CPP_SRCS)) $(PSLP_OBJS)

all: $(LIB)/libhprlp.so $(LIB)/libhprlp.a $(BUILD)/solve_mps_file $(BUILD)/gather_bench $(BUILD)/mps_time

$(BUILD)/%.o: $(SRC)/%.cu $(SRC)/engine.h $(SRC)/kernels.cuh $(SRC)/collective.h $(SRC)/rank_group.h $(SRC)/abi_guard.h include/structs.h include/hprlp_b200.h | $(BUILD)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; false)

$(BUILD)/%.o: $(SRC)/%.cpp $(SRC)/engine.h include/structs.h | $(BUILD)
	$(HOSTCXX) $(CXXFLAGS) -c $< -o $@

$(LIB)/libhprlp.so: $(OBJS) Makefile | $(LIB)
	$(NVCC) -shared $(ARCH) -ccbin $(HOSTCXX) -o $@ $(OBJS) $(LIBS) $(RPATH) -Xlinker -Bsymbolic
	ln -sf libhprlp.so $(LIB)/libhprlp.so.0

$(LIB)/libhprlp.a: $(OBJS) | $(LIB)
	ar rcs $@ $(OBJS)

$(BUILD)/solve_mps_file: $(SRC)/solve_mps_file.cpp $(LIB)/libhprlp.a
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -O2 -Iinclude -o $@ $< $(LIB)/libhprlp.a $(LIBS) $(RPATH)

# measurement tools (not part of the product): gather-ceiling microbenchmark, MPS reader timer
$(BUILD)/gather_bench: tools/gather_bench.cu | $(BUILD)
	$(NVCC) -O3 -std=c++17 $(ARCH) -lineinfo -ccbin $(HOSTCXX) -o $@ $<

$(BUILD)/mps_time: tools/mps_time.cpp $(LIB)/libhprlp.so
	$(HOSTCXX) -O2 -std=c++17 -Iinclude -I$(CUDA_PATH)/include -o $@ $< -L$(LIB) -lhprlp -Wl,-rpath,'$$ORIGIN/../lib'

$(BUILD)/pslp/%.o: $(PSLP_DIR)/src/%.c | $(BUILD)
	@mkdir -p $(dir $@)
	/usr/bin/gcc -O3 -fPIC $(PSLP_INC) -DPSLP_VERSION=\"0.0.8\" -D_POSIX_C_SOURCE=200809L -DNDEBUG -w -c $< -o $@

$(BUILD) $(LIB):
	mkdir -p $@

clean:
	rm -rf $(BUILD) $(LIB)
.PHONY: all clean
