#!/bin/bash
# Round-2 final numbers, v5 code (run-local pair path behind the staged slot order), ONE GPU: the whole -m gpu suite, the
# default bench line, the reference (CPU) arm, the four pair-kernel code paths, then -- each only after its command has
# exited 0 without the profiler -- the ncu launch list of the bench command and one --set full capture of the hot kernels.
O=gpurun_out
timeout 420 python -m pytest tests -m gpu -q > $O/r2_pytest_all_1gpu_v5.log 2>&1; echo "pytest exit $?"; tail -4 $O/r2_pytest_all_1gpu_v5.log
timeout 420 python bench.py > $O/bench_r2_v5_g1.json 2> $O/bench_r2_v5_g1.err; echo "bench exit $?"
timeout 120 python profiles/k1_variants_time.py > $O/k1_variants_time_r2_v5.log 2>&1; echo "variants exit $?"; cat $O/k1_variants_time_r2_v5.log
SFM_REORDER_EVERY=0 timeout 120 python profiles/k1_variants_time.py > $O/k1_variants_time_r2_v5_roworder.log 2>&1
timeout 100 python profiles/kernels_alone.py > $O/kernels_alone_r2_v5.log 2>&1 && \
KA_WARM=0 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k1_sym_pairs|k2_segments" -c 3 -o $O/r2_ncu_full_v5 -f python profiles/kernels_alone.py > $O/ncu_full_v5.log 2>&1; echo "ncu full exit $?"
timeout 200 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-dropin --no-parity > $O/bench_short_v5.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_v5.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-dropin --no-parity > $O/ncu_launches_v5.log 2>&1; echo "ncu launches exit $?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_v5_ref.json 2> $O/bench_r2_v5_ref.err; echo "reference arm exit $?"
python - $O/bench_r2_v5_g1.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    r = d['roofline']
    print('ms/step %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'dropin', (d.get('e2e_dropin') or {}).get('ms_per_step'),
          'k1 alone %.3f' % r['ms_per_launch'], 'frac %.3f' % r['frac'], 'local', r['local_tile_pair_fraction']['timed_ticks'],
          '\nparity', (d.get('parity') or {}).get('oracle'), '\nextra', {k: v.get('ms_per_step') for k, v in d.get('extra', {}).items()},
          '\ncpu', d.get('cpu_baseline'), d['clocks'], 'k2', d['roofline_k2']['ms_alone'])
except Exception as e:
    print('FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-3000:])
PY
