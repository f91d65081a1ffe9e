#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the UNMODIFIED reference CUDA sources, where they lie
# under /root/reference, plus oracle/ref_wrap.cu (our extern "C" wrapper) into
# oracle/_ref/libwast3d_ref.so (git-ignored, shipped to the GPU box by gpurun).
# No reference source is copied into the repo and the reference's own build system
# (setup.py / CMake) is not run.  GCC 13 needs two forced includes (std::uintptr_t, FLT_MAX);
# flags otherwise follow what torch's BuildExtension would pass (nvcc defaults: -O3 device code,
# --fmad=true), arch = sm_100 ("the reference recompiled for this GPU").
set -euo pipefail
REF=${WAST3D_REFERENCE_ROOT:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
DGR="$REF/submodules/diff-gaussian-rasterization"
KNN="$REF/submodules/simple-knn"
if [ ! -d "$DGR/cuda_rasterizer" ]; then
  echo "reference not present at $REF; keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-O3 -std=c++17 -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC
       -include cstdint -include cfloat --expt-relaxed-constexpr -w
       -I"$DGR" -I"$DGR/third_party/glm" -I"$KNN")
pids=()
for src in "$DGR/cuda_rasterizer/forward.cu" "$DGR/cuda_rasterizer/backward.cu" \
           "$DGR/cuda_rasterizer/rasterizer_impl.cu" "$KNN/simple_knn.cu" "$HERE/ref_wrap.cu"; do
  obj="$OUT/$(basename "${src%.cu}").o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT/libwast3d_ref.so" "$OUT"/forward.o "$OUT"/backward.o \
    "$OUT"/rasterizer_impl.o "$OUT"/simple_knn.o "$OUT"/ref_wrap.o -lcudart
echo "built $OUT/libwast3d_ref.so"
