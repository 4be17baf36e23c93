#!/bin/bash
# Build the kernels of another commit into ab/libgsr_b200_<tag>.so for same-box A/B measurements
# (tests/tools/stage_probe.py with GSR_AB_LIB).  usage: build_ab_lib.sh <commit|WORKTREE> <tag> [extra nvcc flags, e.g. -DGSR_AB_FILL_MEMSET]
set -e
commit=$1; tag=$2; shift 2
root=$(cd "$(dirname "$0")/../.." && pwd)
tmp=$(mktemp -d)
if [ "$commit" = "WORKTREE" ]; then
  mkdir -p "$tmp/gs_localization_b200" && cp -r "$root/gs_localization_b200/csrc" "$tmp/gs_localization_b200/" && cp -r "$root/include" "$tmp/"
else
  git -C "$root" archive "$commit" gs_localization_b200/csrc include | tar -x -C "$tmp"
fi
cd "$tmp/gs_localization_b200/csrc"
for f in *.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --cudart shared "$@" -c "$f" -o "$tmp/${f%.cu}.o" &
done
wait
mkdir -p "$root/ab"
nvcc -shared --cudart shared -gencode arch=compute_100a,code=sm_100a -o "$root/ab/libgsr_b200_$tag.so" "$tmp"/*.o -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
rm -rf "$tmp"
echo "$root/ab/libgsr_b200_$tag.so"
