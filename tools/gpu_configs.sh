#!/bin/bash
# One B200: ncu --set full of the first three k_conv3x3_v3 launches of a forward + one bench line per BASELINE config.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_v3 -s 0 -c 3 -o gpurun_out/r02_conv_v3_first \
    python tools/profile_ops.py --precision fp16s --gn-fusion 1 --reps 0 --warmup 1 > gpurun_out/r02_ncu_first.log 2>&1
echo "ncu rc=$?"
for c in 2 3 4 5; do
  steps=40; [ $c = 3 ] && steps=20; [ $c = 5 ] && steps=10
  timeout 1200 python bench.py --config $c --steps $steps --warmup 3 > gpurun_out/r02_bench_config${c}_n1.json 2> gpurun_out/r02_bench_config${c}_n1.err
  echo "config $c rc=$? $(python -c "
import json
d=json.load(open('gpurun_out/r02_bench_config${c}_n1.json'))
print(round(d['value'],3),'img/s', round(d['ms_per_step'],2),'ms/step e2e',round(d['e2e']['value'],3),'roofline',round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None, 'parity', (d['parity_sample'] or {}).get('rel_l2_x0'))
" 2>&1 | tail -1)"
  tail -2 gpurun_out/r02_bench_config${c}_n1.err
done
timeout 600 python bench.py --impl reference --config 2 --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
