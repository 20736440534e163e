#!/bin/bash
# A/B of the once-per-stage loop unrolling (CCC_STAGE_LOOP_UNROLL) on the GPU box: rebuild, time, compare digests.
O=gpurun_out/r02i_ab_srb_gain_prefetch.txt
: > $O
for F in "-DCCC_SRB_GAIN_PREFETCH=0" "-DCCC_SRB_GAIN_PREFETCH=2" "-DCCC_SRB_GAIN_PREFETCH=3" "-DCCC_SRB_GAIN_PREFETCH=5" "-DCCC_SRB_GAIN_PREFETCH=0" "-DCCC_SRB_GAIN_PREFETCH=3"; do
  CCC_EXTRA_NVCC_FLAGS="$F" python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
  python tools/ab_quick.py "flags: $F" srb 0 >> $O 2>&1
done
python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
cat $O
