#!/usr/bin/env bash
# Builds tuning variants of libhprlp.so (different item sizes / occupancy) into lib/variants/ for tools/tune_phase.py.
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p lib/variants
build() {  # name lane_nnz round_nnz min_blocks
  local name=$1; local d=build/var_$name; mkdir -p $d
  local F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xcompiler -fopenmp -Iinclude -DHPR_LANE_NNZ=$2 -DHPR_ROUND_NNZ=$3 -DHPR_MIN_BLOCKS=$4 $5"
  for f in engine api batched; do /usr/local/cuda/bin/nvcc $F -Xptxas -v -c hpr-lp-c_b200/csrc/$f.cu -o $d/$f.o 2> $d/$f.ptxas.log & done
  wait
  /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o lib/variants/libhprlp_$name.so $d/engine.o $d/api.o $d/batched.o build/transpose.o build/partitioned.o build/synth_device.o build/collective.o build/nccl_shim.o build/mps_reader.o build/host_utils.o build/presolve.o build/pslp/core/*.o build/pslp/explorers/*.o \
     -L/usr/local/cuda/lib64 -lcurand -lz -lgomp -lpthread -ldl -Xlinker -rpath -Xlinker /usr/local/cuda/lib64 -Xlinker -Bsymbolic
  echo "$name: $(grep -A3 'XPhaseOpILb0EEELi2E' $d/engine.ptxas.log | grep -E 'Used' | head -1) $(grep -A3 'XPhaseOpILb0EEELi2E' $d/engine.ptxas.log | grep spill | head -1)"
}
for spec in "$@"; do IFS=: read -r n a b c d <<< "$spec"; build $n $a $b $c "$d"; done
