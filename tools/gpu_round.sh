#!/bin/bash
# One GPU session: parity tests, per-kernel sweep, bench line (+ reference arm), launch list, ncu captures of the
# local-correlation kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
tail -3 gpurun_out/tests.log
timeout 600 python tools/bench_kernels.py --variants ${VARIANTS:-0} --no-generic --out gpurun_out/kbench.json > gpurun_out/kbench.log 2>&1
tail -22 gpurun_out/kbench.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.json
if [ -n "$REF_ARM" ]; then timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.json; fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kde4 -o gpurun_out/kde_sym -f python tools/profile_kde.py --pairs 32 --algo 2 > gpurun_out/ncu_kde.log 2>&1
for sc in ${NCU_SCALES:-4 2}; do
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:lc_ -c 3 -o gpurun_out/lc_v2_s$sc -f python tools/profile_lc.py --scale $sc --algo 0 > gpurun_out/ncu_s$sc.log 2>&1
done
ls -la gpurun_out
