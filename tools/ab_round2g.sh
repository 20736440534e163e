#!/bin/bash
# A/B of the once-per-stage loop unrolling (CCC_STAGE_LOOP_UNROLL) on the GPU box: rebuild, time, compare digests.
O=gpurun_out/r02h_ab_srb_unroll_factors.txt
: > $O
for F in "" "-DCCC_SRB_STAGE_UNROLL=2" "-DCCC_SRB_STAGE_UNROLL=4" "-DCCC_SRB_STAGE_UNROLL=1" "-DCCC_TILE_UNROLL=1" "-DCCC_TILE_UNROLL=4" ""; do
  CCC_EXTRA_NVCC_FLAGS="$F" python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
  python tools/ab_quick.py "flags: $F" srb >> $O 2>&1
done
python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
cat $O
