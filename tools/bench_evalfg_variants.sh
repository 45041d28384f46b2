#!/bin/bash
# Build experiment variants of the library on the GPU box and time the config-3 inner loop (wfm_eval_fg) with each.
# usage: tools/bench_evalfg_variants.sh "<name>:<nvcc flags>" ...
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  WFM_BUILD_FLAGS="-DWFM_ONLY_N=512 $flags" python -c "import __graft_entry__ as g; g.build_library(force=True)" > gpurun_out/build_$name.log 2>&1 || { echo "$name: build failed"; tail -5 gpurun_out/build_$name.log; continue; }
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-parity --no-others --sustain 0 > gpurun_out/bench_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{name}.log").read().strip().splitlines()[-1])
    print(f"{name:20s} eval_fg {d['eval_fg']['ms_per_eval']:.4f} ms   step {d['ms_per_step']:.4f} ms")
except Exception as e:
    print(name, "FAILED", e); print(open(f"gpurun_out/bench_{name}.log").read()[-600:])
PY
done
python -c "import __graft_entry__ as g; g.build_library(force=True)" > gpurun_out/build_default.log 2>&1
