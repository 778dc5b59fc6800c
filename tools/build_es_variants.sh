#!/bin/bash
# Build alternative libraries with -DFE_ES_* overrides (tuning aid for the ES forward kernels).
# usage: tools/build_es_variants.sh w12s2 ...        fast kernel: w = warps per block, s = ring stages per warp
#        tools/build_es_variants.sh W12S2U2 ...      streaming kernel: W warps, S stages, U max pairs per warp
# run one with FINENVS_B200_LIB=finenvs_b200/libfe_es_<spec>.so python tools/es_rollout.py ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  if [[ $v =~ ^w([0-9]+)s([0-9]+)$ ]]; then D="-DFE_ES_FAST_WARPS=${BASH_REMATCH[1]} -DFE_ES_FAST_STAGES=${BASH_REMATCH[2]}"
  elif [[ $v =~ ^W([0-9]+)S([0-9]+)U([0-9]+)$ ]]; then D="-DFE_ES_ST_WARPS=${BASH_REMATCH[1]} -DFE_ES_ST_STAGES=${BASH_REMATCH[2]} -DFE_ES_ST_MAXU=${BASH_REMATCH[3]}"
  else echo "bad spec $v"; exit 1; fi
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared \
    -I include $D -o finenvs_b200/libfe_es_$v.so finenvs_b200/csrc/*.cu &
done
wait
ls finenvs_b200/libfe_es_*.so
