#!/bin/bash
# Like bench_variants.sh, but with the sustained (>= 2.5 s) loop of bench.py: burst and power-capped numbers per build.
# usage: tools/bench_sustained_variants.sh "<name>:<nvcc flags>" ...
mkdir -p gpurun_out
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  WFM_BUILD_FLAGS="-DWFM_ONLY_N=512 $flags" python -c "import __graft_entry__ as g; g.build_library(force=True)" > gpurun_out/build_$name.log 2>&1 || { echo "$name: build failed"; continue; }
  timeout 300 python bench.py --steps 50 --warmup 5 --quick --quick-sustain --no-cpu-baseline --no-eval-fg --e2e-steps 1 > gpurun_out/bench_$name.log 2>&1
  python - "$name" <<'PY'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_{name}.log").read().strip().splitlines()[-1])
    s=d.get("roofline_sustained") or {}
    print(f"{name:20s} burst {d['ms_per_step']:.4f} ms/step frac={d['roofline_step']['frac']:.3f} | sustained {s.get('ms_per_step',0):.4f} ms/step frac={s.get('frac',0):.3f} sm={s.get('clocks',{}).get('sm_mhz')}")
except Exception as e:
    print(name, "FAILED", e); print(open(f"gpurun_out/bench_{name}.log").read()[-600:])
PY
done
python -c "import __graft_entry__ as g; g.build_library(force=True)" > gpurun_out/build_default.log 2>&1
