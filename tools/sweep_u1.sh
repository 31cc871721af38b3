# usage: bash tools/sweep_u1.sh "ENV1=a ENV2=b" "ENV1=c" ...   -> kernel_ms of the L=32 bench per environment setting
for cfg in "$@"; do
  echo "CFG: $cfg"
  env $cfg python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('  kernel_ms %.3f  checksum %.12f' % (d['roofline']['kernel_ms'], d['config']['checksum_x_dot_Hx']))
    elif l: print('  ',l[:200])
"
done
