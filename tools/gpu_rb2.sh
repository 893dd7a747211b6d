cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_refiner_blocks.py -m gpu -q 2>&1 | tail -5 > gpurun_out/rb_t1.log
timeout 500 python tools/bench_refiner.py --b 64 > gpurun_out/rb_bench.log 2>&1
timeout 900 python -m pytest tests/test_dropin_reference.py tests/test_library_abi.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/all_gpu.log
cat gpurun_out/rb_t1.log; grep -o '"shape": "[a-z0-9_]*"\|"ms": [0-9.]*\|"dw_ms_b8": [0-9.]*\|"pw_ms_b8": [0-9.]*\|"torch_autocast_ms": [0-9.]*' gpurun_out/rb_bench.log | paste - - - - - ; tail -1 gpurun_out/rb_bench.log; cat gpurun_out/all_gpu.log
