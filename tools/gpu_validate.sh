#!/bin/bash
# Validation of the final build on one B200: full GPU suite, smoke, default bench (no ncu passes - tools/gpu_final.sh has those).
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_final_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_final_gpu.log
grep -E "passed|failed|rc=" gpurun_out/r02_final_gpu.log | tail -4
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_final_smoke.log 2>&1; tail -1 gpurun_out/r02_final_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_config2_n1.json 2> gpurun_out/r02_bench_config2_n1.err; tail -2 gpurun_out/r02_bench_config2_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_config2_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps')}, 'e2e', d['e2e']['value'])
print('roofline', d['roofline']['frac'], d['roofline']['ms_by_kernel_kind'])
print('strict', d['strict_mode']['value'], d['strict_mode']['parity_sample']['rel_l2_x0'])
print('parity', d['parity_sample']['rel_l2_x0'], d['clocks'])
PY
