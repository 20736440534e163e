#!/bin/bash
# A/B of compile-time variants of the DDP kernels on the GPU box: for every set of extra nvcc flags rebuild the library,
# time the fixed workloads of tools/ab_quick.py and print a digest of the outputs (every build must give the same bits).
#   bash tools/ab_flags.sh <output name> [centroidal|srb|""] <flags 1> [<flags 2> ...]
# e.g. bash tools/ab_flags.sh r02x_ab_unroll srb "" "-DCCC_SRB_STAGE_UNROLL=2" "-DCCC_TILE_UNROLL=4" ""
# (repeat the baseline at the end: run-to-run spread is ~2 %, box-to-box ~5 %, so only rows of one file compare).
# Macros: CCC_SRB_STAGE_UNROLL, CCC_TILE_UNROLL, CCC_NO_ROLLED_TILE_LOOPS, CCC_FORCE_ROLLED_TILE_LOOPS (ddp_warp_core.cuh,
# model_srb.cuh), CCC_AB_VARIANTS (ddp_host.cuh).
O=gpurun_out/$1.txt
W=$2
shift 2
mkdir -p gpurun_out
: > $O
for F in "$@"; do
  CCC_EXTRA_NVCC_FLAGS="$F" python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
  python tools/ab_quick.py "flags: $F" "$W" 0 >> $O 2>&1
done
python centroidalcontrolcollection_b200/build.py --force > /dev/null 2>&1
cat $O
