#!/bin/bash
# bench lines at N GPUs (run under `gpurun --gpus N`): tools/gpu_scale.sh N
N=$1; P=29600
for cfg in visir448 map224 map672; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config $cfg --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2_bench_${cfg}_${N}gpu.json
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench_${cfg}_${N}gpu.json')); print('$cfg', d['n_gpus'], 'value %.0f ms %.3f e2e %.0f pipelined %.0f frac %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['pipelined']['value'], d['roofline']['frac']))"
done
