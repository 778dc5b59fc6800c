#!/bin/bash
# Build experiment libraries of the gather kernel: tools/build_gather_variants.sh b6m12s3 b10m8s4 b6m12s3c ...
#   b<book>m<move>s<slots>[c][n]   bookkeeper / mover warp counts; c = with phase clocks (tools/gather_clocks.py); n / l / t = position-feature stores /
#                          gathers / bulk stores switched off (timing experiments; results are wrong)
# run one with FINENVS_B200_LIB=$PWD/finenvs_b200/libfe_ga_<spec>.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
for v in "$@"; do
  if [[ $v =~ ^b([0-9]+)m([0-9]+)s([0-9]+)(c?)(n?)(l?)(t?)(f?)$ ]]; then
    D="-DFE_EXPERIMENTS -DFE_GATHER_BOOK=${BASH_REMATCH[1]} -DFE_GATHER_MOVE=${BASH_REMATCH[2]} -DFE_GATHER_STAGES=${BASH_REMATCH[3]}"
    [[ -n ${BASH_REMATCH[4]} ]] && D="$D -DFE_GATHER_CLOCKS"
    [[ -n ${BASH_REMATCH[5]} ]] && D="$D -DFE_GATHER_NOPF"
    [[ -n ${BASH_REMATCH[6]} ]] && D="$D -DFE_GATHER_NOLOAD"
    [[ -n ${BASH_REMATCH[7]} ]] && D="$D -DFE_GATHER_NOSTORE"
    [[ -n ${BASH_REMATCH[8]} ]] && D="$D -DFE_GATHER_NOFENCE"
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared \
      -I include $D -o finenvs_b200/libfe_ga_$v.so finenvs_b200/csrc/*.cu &
  else echo "bad spec $v"; exit 1; fi
done
wait
ls -la finenvs_b200/libfe_ga_*.so
