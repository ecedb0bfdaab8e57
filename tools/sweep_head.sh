#!/bin/bash
# development sweep: head size (MAS_OPT_APPLY_VARIANT, per-mille of the owned banks; -1 = auto) of the apply graph.
#   VARIANTS="-1 150 250" [WORLD=2] tools/sweep_head.sh
W=${WORLD:-1}
for v in ${VARIANTS:--1}; do
  if [ "$W" = 1 ]; then
    python bench.py --steps 200 --warmup 10 --lean --variant $v 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant',$v, round(d['ms_per_step']*1e3,2),'us', round(d['value']))"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $W --steps 200 --warmup 10 --lean --variant $v 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('world',$W,'variant',$v, round(d['ms_per_step']*1e3,2),'us', round(d['value']))"
  fi
done
