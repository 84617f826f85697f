mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -x -q > gpurun_out/r2s_t.log 2>&1; tail -2 gpurun_out/r2s_t.log
for cfg in "new" "nopertap DPIR_CONV_WPERTAP=0" "noplain3 DPIR_CONV_PLAIN_NORES=0" "old DPIR_CONV_WPERTAP=0 DPIR_CONV_PLAIN_NORES=0"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 300 python tools/profile_ops.py --precision fp16s --gn-fusion 1 > gpurun_out/r2s_ops_$name.txt 2>&1
  echo "$name: $(grep 'total serialised' gpurun_out/r2s_ops_$name.txt | cut -c1-110)"
  grep -E "\('conv(\+gn)?', 256, (128|256), 128, 3, (0|256)\)|\('conv', 128, 256, 256, 3, 0\)|\('conv', 256, 3, 128" gpurun_out/r2s_ops_$name.txt
done
