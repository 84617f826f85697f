#!/bin/bash
# Evidence pass on one B200 (run under gpurun): per-op tables, ncu launch list of the bench command, ncu --set full of the first
# fused convolutions, parity numbers.  Writes to gpurun_out/; summaries are copied into profiles/ on the build box.
mkdir -p gpurun_out
python tools/profile_ops.py --precision fp16s --gn-fusion 1 > gpurun_out/r02_ops_fused_b32.txt 2>&1
python tools/profile_ops.py --precision fp16s --gn-fusion 0 > gpurun_out/r02_ops_unfused_b32.txt 2>&1
python tools/profile_ops.py --precision fp16s --gn-fusion 4 > gpurun_out/r02_ops_fused_r1producers_b32.txt 2>&1
python tools/profile_ops.py --precision fp16x3 > gpurun_out/r02_ops_fp16x3_b32.txt 2>&1
python tools/profile_ops.py --precision fp16s --model imagenet --batch 8 > gpurun_out/r02_ops_imagenet_b8.txt 2>&1
grep "total serialised" gpurun_out/r02_ops_*.txt
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strict-leg > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
# full capture of the first three k_conv3x3_v3 launches of a forward: in[1] conv1 (fused, no residual), conv2 (fused, residual), in[2] conv1 (plain)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_v3 -s 6 -c 3 -o gpurun_out/r02_conv_v3 \
    python tools/profile_ops.py --precision fp16s --gn-fusion 1 --reps 0 --warmup 3 > gpurun_out/r02_ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:k_attention_tc -s 4 -c 2 -o gpurun_out/r02_attention_tc \
    python tools/profile_ops.py --precision fp16s --gn-fusion 1 --reps 0 --warmup 2 > gpurun_out/r02_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
ls -la gpurun_out | grep r02_
