# usage: bash tools/sweep_multi.sh N "ENV=.. [bench args]" ...  -> one bench line summary per configuration on N GPUs
N=$1; shift
for cfg in "$@"; do
  echo "CFG(N=$N): $cfg"
  envs=""; args=""
  for tok in $cfg; do case "$tok" in *=*) envs="$envs $tok";; *) args="$args $tok";; esac; done
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu-baseline $args 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); c=d['config']; print('  ms_per_step %.3f kernel_ms %.3f  %s' % (d['ms_per_step'], d['roofline']['kernel_ms'], json.dumps(c.get('cached_csr', {}).get('ms_per_matvec'))))
    elif 'rror' in l: print('  ',l[:300])
"
done
