#!/bin/bash
# Build alternative libraries with -DFE_PIPE_* overrides for tools/sweep_libs.sh (tuning aid).
# usage: tools/build_pipe_variants.sh b3m8i0o3r8 b4m8i3o2r8 ...   (b=bookkeeper warps, m=mover warps, i=in stages, o=out stages, r=rows/thread)
cd "$(dirname "$0")/.."
for v in "$@"; do
  [[ $v =~ ^b([0-9]+)m([0-9]+)i([0-9]+)o([0-9]+)r([0-9]+)$ ]] || { echo "bad spec $v"; exit 1; }
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared \
    -I include -DFE_PIPE_BOOK=${BASH_REMATCH[1]} -DFE_PIPE_MOVE=${BASH_REMATCH[2]} -DFE_PIPE_SIN=${BASH_REMATCH[3]} \
    -DFE_PIPE_SOUT=${BASH_REMATCH[4]} -DFE_PIPE_RPT=${BASH_REMATCH[5]} -o finenvs_b200/libfe_$v.so finenvs_b200/csrc/*.cu &
done
wait
ls finenvs_b200/libfe_*.so
