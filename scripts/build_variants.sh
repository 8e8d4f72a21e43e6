#!/bin/bash
# Build alternative libraries under wendy_b200/variants/ for scripts/ab_variants.py (they travel to the GPU box
# with the snapshot; the directory is git-ignored).  Usage:
#   scripts/build_variants.sh name1:"-DMACRO=1 ..." name2:"..."      (no arguments: the prepared candidates)
# then e.g.  gpurun -- 'bash scripts/ab_call.sh ab wendy_b200/variants/lib_base.so wendy_b200/variants/lib_*.so'
cd "$(dirname "$0")/../wendy_b200/csrc" || exit 1
mkdir -p ../variants
B="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -fopenmp"
if [ $# -eq 0 ]; then
  set -- "base:" "exact:-DWENDY_FORCE_EXACT_SCAN=1" "wp:-DTK_SLOT_WARPPATH=1" "nw:-DTK_DEST_NOWIN=1" "s32:-DTK_STORE32=1" \
         "all3:-DTK_SLOT_WARPPATH=1 -DTK_DEST_NOWIN=1 -DTK_STORE32=1" "c1024:-DTK_COARSE_CAP=1024" \
         "pe8:-DTK_PERSIST_E=8" "os:-DTK_OWNER_SORT=1" "os8:-DTK_OWNER_SORT=1 -DTK_PERSIST_E=8" \
         "rd:-DTK_ROUNDS=1" "rdos:-DTK_ROUNDS=1 -DTK_OWNER_SORT=1"
fi
for v in "$@"; do
  n=${v%%:*}; f=${v#*:}
  ( $B $f -shared -o ../variants/lib_$n.so api.cu tile.cu wstep.cu small.cu potential.cu radix.cu -lcudart -lgomp \
      > /tmp/build_variant_$n.log 2>&1 && echo "built lib_$n.so ($f)" || { echo "FAILED $n"; tail -5 /tmp/build_variant_$n.log; } ) &
done
wait
