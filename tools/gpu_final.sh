#!/bin/bash
# Final round-2 artefacts with the refiner tail (run on the GPU box; results land in gpurun_out/).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_visir448_1gpu.json 2> gpurun_out/final_bench.err
timeout 500 python tools/bench_refiner.py --b 64 --out gpurun_out/r2_refiner_bench.json > gpurun_out/rb_bench.log 2>&1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py rb_tail_c177 rb_tail_c73 rb_tail_c24 rb_tail_c24_tc flow_update upsample > gpurun_out/r2_sanitizer_refiner_$tool.log 2>&1
  tail -3 gpurun_out/r2_sanitizer_refiner_$tool.log
done
for cfg in "rb_dw p1_s1 dw_s1" "rb_pw_mma p1_s1 pwmma_s1" "rb_dw p1_s16 dw_s16" "rb_pw_kernel p1_s16 pw_s16" "rb_dw p1_s2 dw_s2" "rb_pw_mma p1_s2 pwmma_s2" "rb_pw_kernel p1_s4 pw_s4"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$1 --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_full_$3 python tools/profile_refiner.py --shape $2 --b 64 > gpurun_out/ncu_$3.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_refiner_tail_s2.csv python tools/profile_refiner.py --shape p1_s2 --b 64 > /dev/null 2>&1
cat gpurun_out/final_gpu_tests.log gpurun_out/final_smoke.log; tail -2 gpurun_out/rb_bench.log | head -1 | cut -c1-300; head -c 600 gpurun_out/r2_bench_visir448_1gpu.json; tail -3 gpurun_out/final_bench.err
