mkdir -p gpurun_out
run() { # name, env..., args
  name=$1; shift
  env "$@" timeout 200 python tools/profile_ops.py --precision fp16s --reps 3 --gn-fusion $F > gpurun_out/r2d_$name.txt 2>&1
  echo "$name: $(grep 'total serialised' gpurun_out/r2d_$name.txt | cut -c1-120)"
  grep -E "\('conv(\+gn)?', 256, 128, 128, 3, 0\)|\('conv(\+gn)?', 128, 256, 256, 3, 0\)|\('conv(\+gn)?', 256, 256, 128, 3, 0\)" gpurun_out/r2d_$name.txt
}
F=0 run plain_223 DPIR_CONV_RING=2,2,3
F=0 run plain_341 DPIR_CONV_RING=3,4,1
F=3 run f3_pf0 DPIR_CONV_RING=2,2,3,3,4,1,0
F=3 run f3_pf1 DPIR_CONV_RING=2,2,3,3,4,1,1
F=3 run f3_pf2 DPIR_CONV_RING=2,2,3,3,4,1,2
F=3 run f3_pf3 DPIR_CONV_RING=2,2,3,3,4,1,3
F=3 run f3_pf4 DPIR_CONV_RING=2,2,3,3,4,1,4
F=1 run f1_pf2 DPIR_CONV_RING=2,2,3,3,4,1,2
timeout 900 python -m pytest tests/test_gpu_dropin.py -x -q > gpurun_out/r2d_dropin.log 2>&1; tail -15 gpurun_out/r2d_dropin.log
export DPIR_TEST_EXPERIMENTAL=1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -x -q > gpurun_out/r2d_t.log 2>&1; tail -3 gpurun_out/r2d_t.log
