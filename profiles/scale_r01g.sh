# weak-scaling sweep of the headline metric (run on an 8-GPU box): bash profiles/scale_r01g.sh "8 4 2"
for N in ${1:-8 4 2}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --md-deals 0 --full-games 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'G upd/s', round(d['value']/1e9,2), 'ms/step', round(d['ms_per_step'],4), 'collective', d['collective'], 'env', round(d['env']['value']/1e9,1))"
done
if [ -n "$2" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --md-deals 0 --full-games 0 --collective nccl 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N', d['n_gpus'], 'nccl: G upd/s', round(d['value']/1e9,2), 'ms/step', round(d['ms_per_step'],4))"
fi
