#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own octree-render sources, from where they lie under
# $REF (default /root/reference), into oracle/_ref/:
#   libref_render.so        unmodified reference kernel, rebuilt for sm_100
#   libref_render_instr.so  same sources + a generated visit-log hook (see
#                           oracle/patch_visit_log.py), used for the per-ray
#                           leaf-visit-sequence parity check
# Nothing from $REF is copied into the repository: the instrumented copy lives
# under a temp dir and is deleted after the build; oracle/_ref/ is git-ignored.
# The reference's own build system (CMake + GLEW/GLFW/OpenGL/libpng) is NOT run:
# those dependencies are absent here; the six files below have no GL dependency.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
if [ ! -d "$REF/src/cuda" ]; then
  echo "build_ref: $REF not present — keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/obj_instr"
TORCH_DIR="$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))')"
CXXABI="$($PY -c 'import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')"
INC=(-I"$REF" -I"$REF/3rdparty/cnpy" -I"$REF/3rdparty/glm" -I"$TORCH_DIR/include"
     -I"$TORCH_DIR/include/torch/csrc/api/include" -I/usr/local/cuda/include)
DEFS=(-D_GLIBCXX_USE_CXX11_ABI=$CXXABI)
# The reference sets no arch and no math flags (CMakeLists.txt:1-80): nvcc
# defaults -fmad=true, IEEE div/sqrt, accurate expf.  Only the arch is added.
NVCCFLAGS=(-std=c++17 -O3 --expt-relaxed-constexpr -gencode arch=compute_100,code=sm_100
           -Xcompiler -fPIC -lineinfo -w)
CXXFLAGS=(-std=c++17 -O3 -fPIC -w)
LIBS=(-L"$TORCH_DIR/lib" -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda
      -L/usr/local/cuda/lib64 -lcudart -lz -Wl,-rpath,"$TORCH_DIR/lib")

build_variant() {  # $1 = source root, $2 = obj dir, $3 = output .so, $4.. = extra defs
  local SRC="$1" OBJ="$2" SO="$3"; shift 3
  local EXTRA=("$@")
  local VINC=(-I"$SRC" "${INC[@]:1}")
  # freshness is judged against the reference's own files (the instrumented copy is regenerated
  # on every run, so its mtimes say nothing) and the patch script
  if [ ! -f "$OBJ/renderer_kernel.o" ] || [ "$REF/src/cuda/renderer_kernel.cu" -nt "$OBJ/renderer_kernel.o" ] \
     || [ "$REF/include/cuda/rt_core.cuh" -nt "$OBJ/renderer_kernel.o" ] \
     || [ "$HERE/patch_visit_log.py" -nt "$OBJ/renderer_kernel.o" -a "$SO" != "$OUT/libref_render.so" ]; then
    echo "[build_ref] nvcc renderer_kernel.cu -> $OBJ (takes ~3 min)"
    nvcc "${NVCCFLAGS[@]}" "${VINC[@]}" "${DEFS[@]}" "${EXTRA[@]}" -c "$SRC/src/cuda/renderer_kernel.cu" -o "$OBJ/renderer_kernel.o"
  fi
  [ -f "$OBJ/common.o" ] || nvcc "${NVCCFLAGS[@]}" "${VINC[@]}" "${DEFS[@]}" -c "$REF/src/cuda/common.cu" -o "$OBJ/common.o"
  for f in src/n3tree/n3tree.cpp src/camera.cpp src/data_format.cpp 3rdparty/cnpy/cnpy.cpp; do
    o="$OBJ/$(basename "${f%.cpp}").o"
    [ -f "$o" ] || g++ "${CXXFLAGS[@]}" "${VINC[@]}" "${DEFS[@]}" -c "$REF/$f" -o "$o"
  done
  g++ "${CXXFLAGS[@]}" "${VINC[@]}" "${DEFS[@]}" "${EXTRA[@]}" -c "$HERE/ref_driver.cpp" -o "$OBJ/ref_driver.o"
  g++ -shared -o "$SO" "$OBJ"/*.o "${LIBS[@]}"
  echo "[build_ref] built $SO"
}

WHAT="${1:-all}"
if [ "$WHAT" = all ] || [ "$WHAT" = plain ]; then
  build_variant "$REF" "$OUT/obj" "$OUT/libref_render.so"
fi
if [ "$WHAT" = all ] || [ "$WHAT" = instr ]; then
  TMP="$(mktemp -d)"
  trap 'rm -rf "$TMP"' EXIT
  mkdir -p "$TMP/src" "$TMP/include"
  cp -r "$REF/include/." "$TMP/include/"
  cp -r "$REF/src/cuda" "$TMP/src/cuda"
  $PY "$HERE/patch_visit_log.py" "$TMP"
  build_variant "$TMP" "$OUT/obj_instr" "$OUT/libref_render_instr.so" -DREF_VISIT_LOG
fi
