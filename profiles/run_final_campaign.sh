#!/bin/bash
# usage: run_final_campaign.sh G  -- final-code numbers at G GPUs: weak-ladder point (peer transport) and cfg4
G=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
O=gpurun_out
timeout 300 $TR --master-port 29721 bench.py --gpus $G --steps 20 --warmup 3 > $O/scale_r1_v4_g${G}.json 2> $O/scale_v4_g${G}.err
timeout 300 $TR --master-port 29723 bench.py --gpus $G --workload cfg4 --steps 10 --warmup 3 > $O/cfg4_r1_v4_g${G}.json 2> $O/cfg4_v4_g${G}.err
for f in scale_r1_v4_g${G} cfg4_r1_v4_g${G}; do python - $O/$f.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], d['n_gpus'], d['config']['n_pedestrians'], 'value %.4e' % d['value'], 'ms/step %.3f' % d['ms_per_step'],
          d['kernel_ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], d['ms_per_step_rank0'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
