#!/bin/bash
# usage: run_final_campaign_r2.sh G -- round-2 final-code numbers at G GPUs (one gpurun --gpus G call): the default bench line
# (weak ladder, parity block, cfg5 / cfg4 extras) and, at G = 8, cfg5 as BASELINE.json states it (1,000 ticks) plus the
# 8-rank half of the trajectory agreement check.  Every command is bounded by its own timeout.
G=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
O=gpurun_out
timeout 300 $TR --master-port 29741 bench.py --gpus $G > $O/bench_r2_v4_g${G}.json 2> $O/bench_r2_v4_g${G}.err
if [ "$G" = "8" ]; then
  timeout 600 $TR --master-port 29742 bench.py --gpus 8 --workload cfg5 --steps 1000 --warmup 3 > $O/cfg5_r2_v4_g8_1000steps.json 2> $O/cfg5_r2_v4_g8.err
  timeout 300 $TR --master-port 29743 profiles/cfg5_agreement.py --phase multi --steps 100 > $O/cfg5_agreement_multi_r2_v4.log 2>&1
  tail -2 $O/cfg5_agreement_multi_r2_v4.log
fi
for f in bench_r2_v4_g${G} cfg5_r2_v4_g8_1000steps; do [ -f $O/$f.json ] || continue; python - $O/$f.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1], d['n_gpus'], d['config']['n_pedestrians'], 'value %.4e' % d['value'], 'ms/step %.3f' % d['ms_per_step'],
          'e2e %.3f' % d['e2e']['ms_per_step'], 'k1 alone %.3f' % d['roofline']['ms_per_launch'], 'parity', d.get('parity', {}).get('ok'),
          {k: v.get('ms_per_step') for k, v in d.get('extra', {}).items()}, d['clocks'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
