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
CXXFLAGS := -O2 -std=c++17 -fPIC -Wall -Iinclude -I$(CUDA_PATH)/include
LIBS := -L$(CUDA_PATH)/lib64 -lcurand -lz -lgomp -lpthread -ldl
RPATH := -Xlinker -rpath -Xlinker $(CUDA_PATH)/lib64

CU_SRCS := $(SRC)/engine.cu $(SRC)/api.cu $(SRC)/batched.cu $(SRC)/transpose.cu $(SRC)/partitioned.cu
CPP_SRCS := $(SRC)/mps_reader.cpp $(SRC)/presolve.cpp $(SRC)/nccl_shim.cpp
OBJS := $(patsubst $(SRC)/%.cu,$(BUILD)/%.o,$(CU_SRCS)) $(patsubst $(SRC)/%.cpp,$(BUILD)/%.o,$(CPP_SRCS))

all: $(LIB)/libhprlp.so $(LIB)/libhprlp.a $(BUILD)/solve_mps_file

$(BUILD)/%.o: $(SRC)/%.cu $(SRC)/engine.h $(SRC)/kernels.cuh include/structs.h include/hprlp_b200.h | $(BUILD)
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

$(BUILD) $(LIB):
	mkdir -p $@

clean:
	rm -rf $(BUILD) $(LIB)
.PHONY: all clean
