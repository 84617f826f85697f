#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_comm.py -x -q > gpurun_out/r2w_comm.log 2>&1; echo "comm rc=$?" >> gpurun_out/r2w_comm.log; tail -5 gpurun_out/r2w_comm.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 4 --steps 20 --warmup 3 --weight-bcast library > gpurun_out/r02_bench_config4_n2_libbcast.json 2> gpurun_out/r2w_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2w_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_config4_n2_libbcast.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','weight_broadcast_ms','weight_broadcast_by','rank_consistency')})
PY
