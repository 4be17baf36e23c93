#!/usr/bin/env bash
# TEST INFRASTRUCTURE — not product code.
#
# Builds the UNMODIFIED reference rasterizer (the depth+alpha fork of
# diff-gaussian-rasterization vendored by RPL-CS-UCL/gs_localization) for
# sm_100a, straight from the sources where they lie under /root/reference, and
# "installs" it into oracle/_ref/ exactly as `pip install` of that submodule
# would (one _C extension + the package's __init__.py).  Nothing here is copied
# into git history: oracle/_ref/ is git-ignored (it still travels to the GPU
# box with gpurun).  It is the reference's own build minus its build system:
# direct nvcc on the five translation units named in the submodule's setup.py
# (setup.py:21-29), same default flags (fmad on, no fast-math), plus the one
# flag needed for a modern libstdc++:  -include cstdint
# (rasterizer_impl.h:24 uses std::uintptr_t without including <cstdint>).
#
# Used as: (A) bit-exact oracle for binning / sort / ranges / n_contrib on the
# GPU box, (B) the reference arm of bench.py (`--impl reference`).
set -euo pipefail

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GSR_REFERENCE_ROOT:-/root/reference}/gaussian_splatting/submodules/diff-gaussian-rasterization"
OUT="$HERE/_ref"
PKG="$OUT/diff_gaussian_rasterization"
OBJ="$OUT/obj"

KNN="${GSR_REFERENCE_ROOT:-/root/reference}/gaussian_splatting/submodules/simple-knn"
KPKG="$OUT/simple_knn"

# simple-knn (distCUDA2: mean squared distance to the 3 nearest neighbours, used by create_from_pcd):
# three translation units (its setup.py), plus -include cfloat for FLT_MAX on a modern toolchain.
build_knn() {
  [ -d "$KNN" ] || return 0
  if [ -f "$KPKG/_C.so" ] && [ "$KPKG/_C.so" -nt "$KNN/simple_knn.cu" ] && [ "${1:-}" != "--force" ]; then
    echo "build_ref.sh: $KPKG/_C.so up to date"; return 0
  fi
  mkdir -p "$KPKG" "$OBJ/knn"
  local PY=python TORCH_DIR PYINC
  TORCH_DIR="$($PY -c 'import os, torch; print(os.path.dirname(torch.__file__))')"
  PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
  local FLAGS=(-std=c++17 -O3 --expt-relaxed-constexpr -Xcompiler -fPIC -include cstdint -include cfloat
               -gencode arch=compute_100a,code=sm_100a -lineinfo
               -DTORCH_EXTENSION_NAME=_C -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1 -I"$KNN"
               -isystem "$TORCH_DIR/include" -isystem "$TORCH_DIR/include/torch/csrc/api/include" -isystem "$PYINC" -w)
  nvcc "${FLAGS[@]}" -c "$KNN/simple_knn.cu" -o "$OBJ/knn/simple_knn.o" &
  nvcc "${FLAGS[@]}" -c "$KNN/spatial.cu" -o "$OBJ/knn/spatial.o" &
  nvcc "${FLAGS[@]}" -x cu -c "$KNN/ext.cpp" -o "$OBJ/knn/ext.o" &
  wait
  nvcc -shared -o "$KPKG/_C.so" "$OBJ/knn/simple_knn.o" "$OBJ/knn/spatial.o" "$OBJ/knn/ext.o" \
       -L"$TORCH_DIR/lib" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python \
       -Xlinker -rpath -Xlinker "$TORCH_DIR/lib"
  : > "$KPKG/__init__.py"
  cuobjdump -sass "$OBJ/knn/simple_knn.o" > "$OUT/simple_knn.sass" 2>/dev/null || true
  echo "build_ref.sh: built $KPKG/_C.so"
}

if [ ! -d "$REF" ]; then
  echo "build_ref.sh: reference sources not present at $REF — keeping any prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
build_knn "${1:-}"
if [ -f "$PKG/_C.so" ] && [ "$PKG/_C.so" -nt "$REF/cuda_rasterizer/backward.cu" ] && [ "${1:-}" != "--force" ]; then
  echo "build_ref.sh: $PKG/_C.so up to date"
  exit 0
fi

mkdir -p "$PKG" "$OBJ"
PY=python
TORCH_DIR="$($PY - <<'EOF'
import os, torch
print(os.path.dirname(torch.__file__))
EOF
)"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"

COMMON=(-std=c++17 -O3 --expt-relaxed-constexpr -Xcompiler -fPIC -include cstdint
        -gencode arch=compute_100a,code=sm_100a -lineinfo
        -DTORCH_EXTENSION_NAME=_C -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1
        -I"$REF" -I"$REF/third_party/glm"
        -isystem "$TORCH_DIR/include" -isystem "$TORCH_DIR/include/torch/csrc/api/include"
        -isystem "$PYINC" -w)

pids=()
for src in cuda_rasterizer/rasterizer_impl.cu cuda_rasterizer/forward.cu cuda_rasterizer/backward.cu rasterize_points.cu; do
  o="$OBJ/$(basename "${src%.cu}").o"
  nvcc "${COMMON[@]}" -c "$REF/$src" -o "$o" &
  pids+=($!)
done
nvcc "${COMMON[@]}" -x cu -c "$REF/ext.cpp" -o "$OBJ/ext.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done

nvcc -shared -o "$PKG/_C.so" "$OBJ"/rasterizer_impl.o "$OBJ"/forward.o "$OBJ"/backward.o \
     "$OBJ"/rasterize_points.o "$OBJ"/ext.o \
     -L"$TORCH_DIR/lib" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python \
     -Xlinker -rpath -Xlinker "$TORCH_DIR/lib"

# the "pip install" half: the package's Python front-end, byte-for-byte
install -m 0644 "$REF/diff_gaussian_rasterization/__init__.py" "$PKG/__init__.py"
# keep the forward.cu object's SASS next to it: the binning-critical rounding
# sequences our kernels must reproduce are read off this listing
cuobjdump -sass "$OBJ/forward.o" > "$OUT/forward.sass" 2>/dev/null || true
echo "build_ref.sh: built $PKG/_C.so"
