#!/bin/bash
# what the driver runs at round end, in one go: GPU tests, smoke, the default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/bench_default.json"))
print({k: l[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "gpu_launches")})
print("e2e", l["e2e"]["value"], "frac", l["roofline"]["frac"], "clocks", l["clocks"]["sm_mhz"], l["clocks"]["samples"], l["clocks"]["reasons"])
PY
