#!/bin/bash
# A/B of the once-per-stage loop unrolling (CCC_STAGE_LOOP_UNROLL) on the GPU box: rebuild, time, compare digests.
O=gpurun_out/r02i_ab_fused_rollouts.txt
: > $O
for F in "" "-DCCC_CENTROIDAL_LS_FUSE=2 -DCCC_SRB_LS_FUSE=4"; do
  CCC_EXTRA_NVCC_FLAGS="$F" python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
  for v in 0 2 0 2; do
    python tools/ab_quick.py "flags: $F" "" $v >> $O 2>&1
  done
done
python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
cat $O
