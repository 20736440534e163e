#!/bin/bash
# A/B of the once-per-stage loop unrolling (CCC_STAGE_LOOP_UNROLL) on the GPU box: rebuild, time, compare digests.
O=gpurun_out/r02g_ab_rolled_tile_centroidal.txt
: > $O
# baseline: the library built from the previous commit (copied in by hand; skipped when absent)
if [ -f centroidalcontrolcollection_b200/libccc_b200_prev.so.keep ]; then
  cp centroidalcontrolcollection_b200/libccc_b200.so /tmp/cur.so
  cp centroidalcontrolcollection_b200/libccc_b200_prev.so.keep centroidalcontrolcollection_b200/libccc_b200.so
  python tools/ab_quick.py "previous commit" >> $O 2>&1
  cp /tmp/cur.so centroidalcontrolcollection_b200/libccc_b200.so
fi
for F in "" "-DCCC_FORCE_ROLLED_TILE_LOOPS"; do
  CCC_EXTRA_NVCC_FLAGS="$F" python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
  python tools/ab_quick.py "flags: $F" >> $O 2>&1
done
python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
cat $O
