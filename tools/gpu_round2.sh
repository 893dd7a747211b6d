#!/bin/bash
# Round-2 measurement artefacts (run on the GPU box; results land in gpurun_out/, copied to profiles/ by hand).
set -x
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r2_sanitizer_$tool.log
done
# bench lines (default workload with the CPU baseline, the other BASELINE workloads, the reference arm)
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_visir448_1gpu.json 2> gpurun_out/r2_bench.err
python bench.py --config map224 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_map224_1gpu.json 2>> gpurun_out/r2_bench.err
python bench.py --config map672 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_map672_1gpu.json 2>> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
# launch list + DRAM traffic of one step
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_B32_itr2.csv python tools/profile_step.py --pairs 32 --itr 2 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/lc_traffic.csv python tools/profile_step.py --pairs 32 --itr 2 > /dev/null 2>&1
python tools/lc_traffic.py gpurun_out/lc_traffic.csv gpurun_out/r2_lc_dram_traffic.json > /dev/null
# ncu --set full of the hot kernels inside the step: rot (scale 2), tc2 (scale 8), mma (scale 4), assemble, kde, global match
for k in lc_rot lc_tc2 lc_mma refiner_assemble kde4_sym gm_tc; do
  skip=0; [ $k = lc_tc2 ] && skip=2; [ $k = refiner_assemble ] && skip=6
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k --launch-skip $skip --launch-count 1 -f -o gpurun_out/r2_full_$k python tools/profile_step.py --pairs 32 --itr 2 > gpurun_out/ncu_full_$k.log 2>&1
done
