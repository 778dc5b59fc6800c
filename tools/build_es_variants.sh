#!/bin/bash
# Build alternative libraries with -DFE_ES_FAST_* overrides (tuning aid for fe_es_forward_fast_kernel).
# usage: tools/build_es_variants.sh w8s2 w12s2 ...   (w = warps per block, s = ring stages per warp)
# run one with FINENVS_B200_LIB=finenvs_b200/libfe_es_<spec>.so python tools/es_rollout.py ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  [[ $v =~ ^w([0-9]+)s([0-9]+)$ ]] || { echo "bad spec $v"; exit 1; }
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared \
    -I include -DFE_ES_FAST_WARPS=${BASH_REMATCH[1]} -DFE_ES_FAST_STAGES=${BASH_REMATCH[2]} \
    -o finenvs_b200/libfe_es_$v.so finenvs_b200/csrc/*.cu &
done
wait
ls finenvs_b200/libfe_es_*.so
