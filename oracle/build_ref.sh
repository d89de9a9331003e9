#!/usr/bin/env bash
# TEST/BENCH INFRASTRUCTURE ONLY -- never linked into the product library.
#
# Builds the UNMODIFIED reference (PolyU-IOR/HPR-LP-C) from the sources where they lie
# under /root/reference into oracle/_ref/ (git-ignored, shipped to the GPU box by gpurun):
#   oracle/_ref/libhprlp_ref.so   the reference's own CUDA build (sm_100), same 7 C symbols
#   oracle/_ref/solve_mps_file    the reference CLI
# This is our own short recipe (nvcc/gcc on the source files directly); the reference's
# Makefile/CMake are not run and no reference source is copied into the repo.
# The reference has NO CPU path, so this CUDA build is the baseline and the strongest oracle
# (SURVEY.md 8c); it can only execute on the GPU box.
set -euo pipefail
REF=${HPRLP_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "[build_ref] $REF not present; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj/pslp"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100,code=sm_100"
NVFLAGS="-w -O2 --std=c++11 $ARCH -Xcompiler -fPIC -Xcompiler -D_GLIBCXX_USE_CXX11_ABI=0"
INC="-I$REF/include -I$REF/include/cuda_kernels"
P=$REF/third_party/PSLP
PINC="-I$P/include/PSLP -I$P/include/core -I$P/include/data_structures -I$P/include/explorers"
PDEF=(-DPSLP_VERSION=\"0.0.8\" -D_POSIX_C_SOURCE=200809L -DNDEBUG)

pids=()
for f in pslp_integration.cpp mps_reader.cpp utils.cu scaling.cu preprocess.cu power_iteration.cu \
         main_iterate.cu HPRLP.cu batched_solver.cu cuda_kernels/HPR_cuda_kernels.cu; do
  o="$OUT/obj/$(basename "${f%.*}").o"
  if [ ! -f "$o" ] || [ "$REF/src/$f" -nt "$o" ]; then
    $NVCC $NVFLAGS $INC $PINC -c "$REF/src/$f" -o "$o" &
    pids+=($!)
  fi
done
for f in $(ls $P/src/core/*.c $P/src/explorers/*.c | grep -v Debugger.c); do
  o="$OUT/obj/pslp/$(basename "${f%.c}").o"
  if [ ! -f "$o" ]; then
    gcc -O3 -fPIC $PINC "${PDEF[@]}" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done

$NVCC -shared $NVFLAGS -o "$OUT/libhprlp_ref.so" $OUT/obj/*.o $OUT/obj/pslp/*.o \
  -L/usr/local/cuda/lib64 -lcublas -lcusparse -lcurand -lz \
  -Xlinker --exclude-libs,ALL -Xcompiler -static-libstdc++ -Xcompiler -static-libgcc
ar rcs "$OUT/libhprlp_ref.a" $OUT/obj/*.o $OUT/obj/pslp/*.o
$NVCC $NVFLAGS $INC -o "$OUT/solve_mps_file" "$REF/src/solve_mps_file.cpp" "$OUT/libhprlp_ref.a" \
  -L/usr/local/cuda/lib64 -lcublas -lcusparse -lcurand -lz
cp "$REF/data/model.mps" "$OUT/model.mps"   # data fixture used by config 1 (not source code)
rm -f "$OUT/libhprlp_ref.a"
echo "[build_ref] built $OUT/libhprlp_ref.so and $OUT/solve_mps_file"
