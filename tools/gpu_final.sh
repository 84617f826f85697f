#!/bin/bash
# Final validation + evidence on one B200: full GPU suite, smoke, default bench, per-op tables, ncu launch list and full captures.
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
print('roofline', d['roofline']['frac'], d['roofline']['ms_by_kernel_kind'], 'traffic', d['roofline']['traffic'])
print('strict', d['strict_mode']['value'], d['strict_mode']['parity_sample']['rel_l2_x0'])
print('parity', d['parity_sample']['rel_l2_x0'], d['clocks'])
PY
python tools/profile_ops.py --precision fp16s --gn-fusion 1 > gpurun_out/r02_ops_fused_b32.txt 2>&1
python tools/profile_ops.py --precision fp16s --gn-fusion 0 > gpurun_out/r02_ops_unfused_b32.txt 2>&1
python tools/profile_ops.py --precision fp16x3 > gpurun_out/r02_ops_fp16x3_b32.txt 2>&1
python tools/profile_ops.py --precision fp16s --model imagenet --batch 8 > gpurun_out/r02_ops_imagenet_b8.txt 2>&1
grep "total serialised" gpurun_out/r02_ops_fused_b32.txt gpurun_out/r02_ops_unfused_b32.txt gpurun_out/r02_ops_fp16x3_b32.txt gpurun_out/r02_ops_imagenet_b8.txt
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strict-leg > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
# full captures: the first three k_conv3x3_v3 launches of a forward (stem, in[1] conv1 fused, conv2 fused + residual) and the output head
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_v3 -s 0 -c 3 -o gpurun_out/r02_conv_v3 \
    python tools/profile_ops.py --precision fp16s --gn-fusion 1 --reps 0 --warmup 1 > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu v3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_dxn -s 0 -c 1 -o gpurun_out/r02_conv_dxn \
    python tools/profile_ops.py --precision fp16s --gn-fusion 1 --reps 0 --warmup 1 > gpurun_out/r02_ncu_dxn.log 2>&1
echo "ncu dxn rc=$?"
ls -la gpurun_out | grep -E "ncu-rep|launches"
