for lib in "" build/libsfm_k2mb5.so build/libsfm_k2mb6.so build/libsfm_k2mb8.so; do
  echo "== lib=${lib:-default(mb4)}"
  SFM_LIB=${lib:+$PWD/$lib} SFM_OVERLAP=0 python bench.py --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('  no-overlap step', round(d['ms_per_step'],3), d['kernel_ms_per_step'])"
  SFM_LIB=${lib:+$PWD/$lib} python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('  overlap step', round(d['ms_per_step'],3))"
done
