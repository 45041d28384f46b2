#!/bin/bash
# Build experiment variants of the library on the GPU box and bench each (headline size only).
# usage: tools/bench_variants.sh "<name>:<nvcc flags>[:<env assignments>]" ...
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; rest="${spec#*:}"; flags="${rest%%:*}"; envs=""
  if [[ "$rest" == *:* ]]; then envs="${rest#*:}"; fi
  WFM_BUILD_FLAGS="-DWFM_ONLY_N=${BENCH_N:-512} $flags" python -c "import __graft_entry__ as g; g.build_library(force=True)" > gpurun_out/build_$name.log 2>&1 || { echo "$name: build failed"; tail -5 gpurun_out/build_$name.log; continue; }
  env $envs timeout 300 python bench.py --nxy ${BENCH_N:-512} --nz ${BENCH_NZ:-256} ${BENCH_EXTRA:-} --steps 50 --warmup 5 --quick --no-cpu-baseline --no-eval-fg --e2e-steps 1 > gpurun_out/bench_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{name}.log").read().strip().splitlines()[-1])
    print(f"{name:28s} {d['value']:10.0f} planes/s  {d['ms_per_step']:.4f} ms/step  step_frac={d['roofline_step']['frac']:.3f}  k={d['kernel_ms_per_step']}")
except Exception as e:
    print(name, "FAILED", e); print(open(f"gpurun_out/bench_{name}.log").read()[-600:])
PY
done
python -c "import __graft_entry__ as g; g.build_library(force=True)" > gpurun_out/build_default.log 2>&1
