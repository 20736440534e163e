#!/bin/bash
# tools/bench_matrix.sh "<variants>" "<chunks>" [extra bench args]: short bench runs over launch shapes x chunk sizes.
variants=${1:-"0 1 2"}; chunks=${2:-"32 0"}; shift 2
mkdir -p gpurun_out
for v in $variants; do for c in $chunks; do
  python bench.py --steps 2 --warmup 1 --variant $v --chunk $c --no-cpu-baseline "$@" 2>&1 | tail -1 > gpurun_out/bm_${v}_${c}.json
  python - "$v" "$c" gpurun_out/bm_${v}_${c}.json <<'PY'
import json, sys
v, c, f = sys.argv[1:4]
try:
    d = json.load(open(f))
    print(f"variant {v} chunk {c}: {d['value']:.0f} solves/s  {d['ms_per_step']:.1f} ms/step  e2e {d['e2e']['value']:.0f}  "
          f"iters {d['config']['mean_ddp_iters']:.1f}  power {d['clocks']['power_w_max']}")
except Exception as e:
    print(f"variant {v} chunk {c}: FAILED {e}: {open(f).read()[:300]}")
PY
done; done
