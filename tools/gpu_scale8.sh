#!/bin/bash
# 8 B200s of one box: BASELINE configs 3, 4, 5 (strong scaling: fixed total batch) through torchrun, one JSON line each.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29611
for c in 3 4 5; do
  steps=40; [ $c = 5 ] && steps=20
  timeout 600 $TR --master-port $port bench.py --gpus 8 --config $c --steps $steps --warmup 5 --profile-steps 2 \
      > gpurun_out/r02_bench_config${c}_n8.json 2> gpurun_out/r02_bench_config${c}_n8.err
  echo "config $c N=8 rc=$? $(python -c "
import json
d=json.load(open('gpurun_out/r02_bench_config${c}_n8.json'))
print(round(d['value'],3),'img/s', round(d['ms_per_step'],2),'ms/step e2e',round(d['e2e']['value'],3),'roofline',round(d['roofline']['frac'],3),'bcast ms',round(d['weight_broadcast_ms'],2),d['rank_consistency']['bit_identical'],d['config']['per_gpu_batch'])
" 2>&1 | tail -1)"
  tail -2 gpurun_out/r02_bench_config${c}_n8.err
  port=$((port+1))
done
