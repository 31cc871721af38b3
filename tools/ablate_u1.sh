# needs a library built with -DEDCUDA_PROFILING (make -C exactdiagonalization.jl_b200/csrc clean all EXTRA=-DEDCUDA_PROFILING): the ablation knob drops parts of the Hamiltonian and is not compiled into release builds
for a in 0 1 2 4 8 16 32 63 3; do
  echo "ABLATE=$a"; EDCUDA_U1_ABLATE=$a python bench.py --steps 5 --warmup 3 --no-e2e --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('  kernel_ms',d['roofline']['kernel_ms'])
    elif l: print('  ',l[:200])
"
done
