cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_refiner_blocks.py -m gpu -q 2>&1 | tail -3
for cfg in "GFB_DW_PW=8" "GFB_DW_PW=4" "GFB_DW_PW=8 GFB_NO_FUSE=1"; do
  echo "== $cfg"; env $cfg timeout 500 python tools/bench_refiner.py --b 64 --torch 0 --out gpurun_out/rb_try.json | grep -o '"shape": "[a-z0-9_]*"\|"ms": [0-9.]*\|"dw_ms_b8": [0-9.]*\|"sum_ms": [0-9.]*' | paste - - - | tr '\n' ' '; echo
done
