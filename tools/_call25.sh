#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_shapes.py -x -q > gpurun_out/r2x_t.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_t.log; tail -3 gpurun_out/r2x_t.log
timeout 300 python tools/profile_ops.py --precision fp16s --gn-fusion 1 > gpurun_out/r2x_ops_il.txt 2>&1
DPIR_SKIP_INTERLEAVE=0 timeout 300 python tools/profile_ops.py --precision fp16s --gn-fusion 1 > gpurun_out/r2x_ops_noil.txt 2>&1
DPIR_TC_STAGGER=1 timeout 300 python tools/profile_ops.py --precision fp16s --gn-fusion 1 > gpurun_out/r2x_ops_stagger.txt 2>&1
timeout 300 python tools/profile_ops.py --precision fp16s --gn-fusion 1 > gpurun_out/r2x_ops_il2.txt 2>&1
grep "total serialised" gpurun_out/r2x_ops_*.txt
python - <<'PY'
import re
def load(f):
    rows=[]
    for l in open(f):
        m=re.match(r'(\w+)\s+\(([^)]*)\)\s+([\d.]+)',l)
        if m: rows.append((m.group(1),m.group(2),float(m.group(3))))
        if l.startswith('total'): break
    return rows
a=load('gpurun_out/r2x_ops_il.txt')
for name in ('noil','stagger','il2'):
    b=load(f'gpurun_out/r2x_ops_{name}.txt')
    skip=sum(y[2]-x[2] for x,y in zip(a,b) if x[0]=='conv' and not x[1].endswith(', 0'))
    small=sum(y[2]-x[2] for x,y in zip(a,b) if x[0]=='conv' and (x[1].startswith('8, 8') or x[1].endswith('1, 0')))
    rest=sum(y[2]-x[2] for x,y in zip(a,b)) - skip - small
    print(name, 'minus il: skip-convs %+.3f  small/1x1 %+.3f  rest %+.3f' % (skip, small, rest))
PY
