mkdir -p gpurun_out
export DPIR_TEST_EXPERIMENTAL=1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -x -q > gpurun_out/r2_t2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t2.log
tail -5 gpurun_out/r2_t2.log
for f in 0 1 3; do
timeout 300 python tools/profile_ops.py --precision fp16s --gn-fusion $f > gpurun_out/r2b_ops_f$f.txt 2>&1
done
grep "total serialised" gpurun_out/r2b_ops_f*.txt
