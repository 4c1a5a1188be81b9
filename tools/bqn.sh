#!/bin/bash
# quick N-GPU bench line (weak scaling): ms/step + phases + dist_parity.  usage: tools/bqn.sh N [extra bench args]
N=${1:-2}; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=%d ms/step %.3f' % (d['n_gpus'], d['ms_per_step']), d.get('dist_parity',{}).get('worst_rel_err'), {k: round(v,3) for k,v in d['roofline']['phases_ms_per_step'].items()}, {k: round(v,3) for k,v in d['roofline'].get('poisson_ms_per_step',{}).items()})"
