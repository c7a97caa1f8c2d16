#!/bin/bash
# TEST INFRASTRUCTURE: builds the REFERENCE's own GPU detector, from its sources where they lie
# under /root/reference (nothing is copied into this repository), into oracle/_ref/librefgpu.so.
#
#   reference sources : threshold.cu labeling_allegretti_2019_BKE.cu line_fit_filter.cu points.cu
#                       cuda_frc971.cu apriltag_gpu.cu apriltag_detect.cu      (unmodified)
#   shims (oracle/ref_shims): glog / gflags macro stand-ins, libapriltag declarations, and a
#                       random-access TransformOutputIterator (the toolkit's CUB 2.8 needs operator+;
#                       the reference pins CCCL 2.3.2) found first on the include path
#   harness           : oracle/ref_harness.cu (libapriltag entry points + flat C API)
#
# The reference's own build system (colcon/CMake + fetched dependencies) is not run.
# oracle/_ref/ is git-ignored but travels to the GPU box with gpurun.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}/src/apriltags_cuda"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) -- keeping any prebuilt $OUT/librefgpu.so" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr --extended-lambda
       -Xcompiler -fPIC -w
       -I "$HERE/ref_shims" -I "$REF/include" -I "$HERE/../include/apriltag_compat")
SRCS=(threshold labeling_allegretti_2019_BKE line_fit_filter points cuda_frc971 apriltag_gpu apriltag_detect)
pids=()
for s in "${SRCS[@]}"; do
  if [ ! -f "$OUT/obj/$s.o" ] || [ "$REF/src/$s.cu" -nt "$OUT/obj/$s.o" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$REF/src/$s.cu" -o "$OUT/obj/$s.o" &
    pids+=($!)
  fi
done
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_harness.cu" -o "$OUT/obj/ref_harness.o" &
pids+=($!)
gcc -O2 -fPIC -I "$HERE/../include/apriltag_compat" -I "$HERE/../ros_vision_b200/csrc" \
    -c "$HERE/../ros_vision_b200/csrc/apriltag_compat.c" -o "$OUT/obj/apriltag_compat.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
OBJS=()
for s in "${SRCS[@]}"; do OBJS+=("$OUT/obj/$s.o"); done
"$NVCC" -shared -o "$OUT/librefgpu.so" "${OBJS[@]}" "$OUT/obj/ref_harness.o" "$OUT/obj/apriltag_compat.o" \
    -gencode arch=compute_100a,code=sm_100a -lcudart
echo "built $OUT/librefgpu.so"
