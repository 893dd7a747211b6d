cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_refiner_blocks.py -m gpu -v -k "pack or dw5 or out_conv or flow_update or upsample or rejects" 2>&1 | tail -45 > gpurun_out/rb_t1.log
timeout 300 python -m pytest tests/test_refiner_blocks.py -m gpu -v -s -k "pointwise or golden" 2>&1 | grep -E "PASS|FAIL|ERROR|tail \(|assert|Error|passed|failed" | head -60 > gpurun_out/rb_t2.log
timeout 400 python -m pytest tests/test_refiner_blocks.py -m gpu -v -s -k "autocast" 2>&1 | grep -E "PASS|FAIL|ERROR|C=|assert|Error|passed|failed" | head -40 > gpurun_out/rb_t3.log
timeout 500 python tools/bench_refiner.py --b 64 > gpurun_out/rb_bench.log 2>&1
cat gpurun_out/rb_t1.log gpurun_out/rb_t2.log gpurun_out/rb_t3.log; tail -12 gpurun_out/rb_bench.log
