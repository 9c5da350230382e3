#!/bin/bash
# 8-GPU measurement campaign (one gpurun --gpus 8 call): weak ladder (both transports), cfg4, cfg5 (1,000 steps), and the
# 8-rank half of the cfg5 trajectory agreement.  Every command is bounded by its own timeout.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
O=gpurun_out
timeout 300 $TR --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 3 > $O/scale_r1_v3_g8_peer.json 2> $O/scale_g8_peer.err
SFM_EXCHANGE=nccl timeout 300 $TR --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 3 > $O/scale_r1_v3_g8_nccl.json 2> $O/scale_g8_nccl.err
timeout 300 $TR --master-port 29703 bench.py --gpus 8 --workload cfg4 --steps 20 --warmup 3 > $O/cfg4_r1_v3_g8.json 2> $O/cfg4_g8.err
timeout 600 $TR --master-port 29704 bench.py --gpus 8 --workload cfg5 --steps 1000 --warmup 3 > $O/cfg5_r1_v3_g8_1000steps.json 2> $O/cfg5_g8.err
timeout 400 $TR --master-port 29705 profiles/cfg5_agreement.py --phase multi --steps 100 > $O/cfg5_agreement_multi.log 2>&1
for f in scale_r1_v3_g8_peer scale_r1_v3_g8_nccl cfg4_r1_v3_g8 cfg5_r1_v3_g8_1000steps; do python - $O/$f.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], d['n_gpus'], d['config']['n_pedestrians'], 'value %.4e' % d['value'], 'ms/step %.3f' % d['ms_per_step'],
          d['kernel_ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
tail -2 $O/cfg5_agreement_multi.log
