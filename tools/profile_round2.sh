#!/bin/bash
# Round-2 ncu captures of every kernel outside the DDP headline kernel (run on the GPU box, one GPU):
#   bash tools/profile_round2.sh     -> gpurun_out/r02_*.ncu-rep + text summaries
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -f"
# QP engine, packed-R kernel (two CTAs per SM): config 2 (26 iterations per QP) and config 5 (5 iterations)
$NCU -k regex:qp_solve_kernel -c 1 -o $O/r02_qp_cfg2 python tools/profile_qp.py 2 592 0 > $O/r02_qp_cfg2.log 2>&1
$NCU -k regex:qp_solve_kernel -c 1 -o $O/r02_qp_cfg5 python tools/profile_qp.py 5 4096 0 > $O/r02_qp_cfg5.log 2>&1
# LinearMpcXY device pipeline: condense, hessian (DMMA), gradient, QP setup, QP solve (256 threads, matrix groups)
$NCU -k regex:'xy_|qp_setup|qp_solve' -c 5 -o $O/r02_xy python tools/profile_xy.py 148 8 --once > $O/r02_xy.log 2>&1
# DdpZmp, one thread per problem
$NCU -k regex:zmp_thread_kernel -c 1 -o $O/r02_zmp_thread python tools/profile_zmp.py 65536 3 1 --once > $O/r02_zmp_thread.log 2>&1
# schedule compiler + device-side ISMPC planOnce (assembly, post) on config 5's 256 plans x 64 perturbations
$NCU -k regex:'footstep_compile|zmp_assemble|zmp_post' -c 3 -o $O/r02_zmp_mpc python -m pytest tests/test_gpu_zmp_mpc.py -q -k config5 > $O/r02_zmp_mpc.log 2>&1
# DdpSingleRigidBody: 1184 problems (one per resident warp), 30 iterations
$NCU -k regex:ddp_solve_kernel -c 1 -o $O/r02_srb python tools/profile_srb.py 1184 30 --once > $O/r02_srb.log 2>&1
for r in r02_qp_cfg2 r02_qp_cfg5 r02_xy r02_zmp_thread r02_zmp_mpc r02_srb; do
  python tools/ncu_kernels.py $O/$r.ncu-rep > $O/${r}_summary.txt 2>&1
done
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
