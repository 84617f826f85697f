mkdir -p gpurun_out
run() { # name, env..., args
  name=$1; shift
  env "$@" timeout 200 python tools/profile_ops.py --precision fp16s --reps 3 --gn-fusion $F > gpurun_out/r2c_$name.txt 2>&1
  echo "$name: $(grep 'total serialised' gpurun_out/r2c_$name.txt | cut -c1-120)"
  grep -E "\('conv(\+gn)?', 256, 128, 128, 3, 0\)|\('conv(\+gn)?', 128, 256, 256, 3, 0\)|\('conv(\+gn)?', 256, 256, 128, 3, 0\)" gpurun_out/r2c_$name.txt
}
F=0 run plain_26 DPIR_CONV_RING=2,6
F=0 run plain_34 DPIR_CONV_RING=3,4
F=0 run plain_25 DPIR_CONV_RING=2,5
F=3 run f3_34 DPIR_CONV_RING=2,6,3,4
F=3 run f3_34_noload DPIR_CONV_RING=2,6,3,4 DPIR_CONV_DEBUG=32
F=3 run f3_34_nomath DPIR_CONV_RING=2,6,3,4 DPIR_CONV_DEBUG=64
F=3 run f3_34_nofence DPIR_CONV_RING=2,6,3,4 DPIR_CONV_DEBUG=128
F=3 run f3_34_none DPIR_CONV_RING=2,6,3,4 DPIR_CONV_DEBUG=224
F=3 run f3_26 DPIR_CONV_RING=2,6,2,6
F=1 run f1_26 DPIR_CONV_RING=2,6,2,6
F=1 run f1_34 DPIR_CONV_RING=2,6,3,4
