#!/bin/bash
# Memory check without a GPU (compute-sanitizer is closed on this pool): the kernel sources are compiled for the CPU
# execution-model emulator WITH AddressSanitizer and a set of MixedOp / cell / RHN / head shapes (multi-segment rows,
# every stride, ragged batches) runs forward + backward through the C ABI.  TEST INFRASTRUCTURE ONLY.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=${TMPDIR:-/tmp}/gnas_asan_build
mkdir -p "$B"
FLAGS="-O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -fPIC -fopenmp -DGNAS_EMUL -x c++ -include $ROOT/tests/emul/cuda_emul.h -Wno-unknown-pragmas -Wno-attributes"
OBJS=""
for f in "$ROOT"/genome-nas_b200/csrc/*.cu "$ROOT"/tests/emul/cuda_emul.cpp; do
  o="$B/$(basename "${f%.*}").o"; g++ $FLAGS -c "$f" -o "$o" & OBJS="$OBJS $o"
done
wait
g++ -shared -fsanitize=address -fopenmp -o "$B/libgnas_emul_asan.so" $OBJS
GNAS_ASAN_LIB="$B/libgnas_emul_asan.so" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 \
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) python "$ROOT/tools/asan_emul_cases.py"
