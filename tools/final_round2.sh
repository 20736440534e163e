#!/bin/bash
# Final single-GPU measurements of round 2 (run on the GPU box): counters, bench line, reference arm, launch list.
export CCC_GIT_REV=${CCC_GIT_REV:-e358007}
O=gpurun_out
mkdir -p $O
python tools/capture_counters.py centroidal 16384 > $O/r02h_counters_centroidal.log 2>&1
python tools/capture_counters.py srb 8192 > $O/r02h_counters_srb.log 2>&1
cp $O/r02_ddp_centroidal_counters.json $O/r02_ddp_srb_counters.json profiles/ 2>/dev/null  # bench.py reads them from profiles/
python bench.py > $O/r02h_bench_1gpu.json 2> $O/r02h_bench_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02h_bench_reference.json 2> $O/r02h_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02h_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-other --no-cpu-baseline > $O/r02h_bench_under_ncu.log 2>&1
tail -c 600 $O/r02h_bench_1gpu.json; tail -3 $O/r02h_bench_1gpu.err; tail -c 400 $O/r02h_bench_reference.json
