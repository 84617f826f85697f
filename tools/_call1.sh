mkdir -p gpurun_out
export DPIR_TEST_EXPERIMENTAL=1
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -k "gn_fus" -x -q > gpurun_out/r2_t_mode3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t_mode3.log
timeout 300 python tools/profile_ops.py --precision fp16s --gn-fusion 1 > gpurun_out/r2_ops_f1.txt 2>&1
timeout 300 python tools/profile_ops.py --precision fp16s --gn-fusion 3 > gpurun_out/r2_ops_f3.txt 2>&1
tail -5 gpurun_out/r2_t_mode3.log
grep "total serialised" gpurun_out/r2_ops_f1.txt gpurun_out/r2_ops_f3.txt
