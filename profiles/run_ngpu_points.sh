#!/bin/bash
# usage: run_ngpu_points.sh G   -- weak-ladder point (both transports), cfg4 and the cfg5 strong-scaling point at G GPUs
G=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
O=gpurun_out
timeout 300 $TR --master-port 29711 bench.py --gpus $G --steps 20 --warmup 3 > $O/scale_r1_v3_g${G}_peer.json 2> $O/scale_g${G}_peer.err
SFM_EXCHANGE=nccl timeout 300 $TR --master-port 29712 bench.py --gpus $G --steps 20 --warmup 3 > $O/scale_r1_v3_g${G}_nccl.json 2> $O/scale_g${G}_nccl.err
timeout 300 $TR --master-port 29713 bench.py --gpus $G --workload cfg4 --steps 10 --warmup 3 > $O/cfg4_r1_v3_g${G}.json 2> $O/cfg4_g${G}.err
timeout 400 $TR --master-port 29714 bench.py --gpus $G --workload cfg5 --steps 5 --warmup 3 > $O/cfg5_r1_v3_g${G}.json 2> $O/cfg5_g${G}.err
for f in scale_r1_v3_g${G}_peer scale_r1_v3_g${G}_nccl cfg4_r1_v3_g${G} cfg5_r1_v3_g${G}; do python - $O/$f.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], d['n_gpus'], d['config']['n_pedestrians'], 'value %.4e' % d['value'], 'ms/step %.3f' % d['ms_per_step'],
          d['kernel_ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
