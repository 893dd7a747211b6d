#!/bin/bash
# Last artefact refresh of the round (1 GPU): tests, smoke, the three workloads, the reference arm, refiner bench.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_visir448_1gpu.json 2> gpurun_out/final_bench.err
timeout 600 python bench.py --config map224 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_map224_1gpu.json 2>> gpurun_out/final_bench.err
timeout 600 python bench.py --config map672 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_map672_1gpu.json 2>> gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/final_bench.err
timeout 600 python tools/bench_refiner.py --b 64 --out gpurun_out/r2_refiner_bench.json > gpurun_out/rb_bench.log 2>&1

cat gpurun_out/final_gpu_tests.log gpurun_out/final_smoke.log; tail -1 gpurun_out/rb_bench.log | cut -c1-400; tail -3 gpurun_out/final_bench.err
for f in visir448 map224 map672; do python - "$f" <<'PY'
import json, sys
l = json.load(open(f"gpurun_out/r2_bench_{sys.argv[1]}_1gpu.json"))
t, fd = l.get("refiner_tail") or {}, l.get("full_decoder") or {}
print(sys.argv[1], round(l["value"], 1), round(l["ms_per_step"], 3), "e2e", round(l["e2e"]["value"], 1), "frac", round(l["roofline"]["frac"], 4),
      "tail_ms", t.get("ms_per_step", t.get("unavailable")), "full", fd.get("pairs_per_s", fd.get("unavailable")), fd.get("decoder_ms_per_step"))
PY
done
