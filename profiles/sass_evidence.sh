#!/bin/bash
# regenerates profiles/r2_sass_evidence.txt (v4) / r2_sass_evidence_v5.txt: opcode counts that show what the kernels are built from (UBLKCP = TMA bulk
# copy, SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed FP32x2, REDG = the integer accumulator atomics, MEMBAR.*.SYS =
# the system-scope fences of the peer barrier).  Run from the repo root after __graft_entry__.build().
LIB=carla-social-force-model_b200/sfm_b200/libsfm_b200.so
for k in k1_sym_pairsILb0ELb0 k2_segmentsILi0 k2_segmentsILi1 k2_msort_scatter k3_integrate k7_barrier k1_sym_finish; do
  echo "## $k"
  cuobjdump -sass $LIB | awk -v k="$k" '/Function :/{f=($0 ~ k)} f' | grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" |
    awk '{print $NF}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -16 | tr '\n' ';'
  echo
done

# instruction mix of the planar inner loop of the pair kernel (one step = 2 x 4 rows = 8 packed calls = 16 pair terms per lane) and of the
# general (3-D) loop: the bodies of the two backward branches with > 150 instructions that hold no scalar FFMA
echo "## k1_sym_pairs<false,false> inner loops (per step of 8 packed calls: 4 rows per thread x 2 j-pairs)"
cuobjdump -sass $LIB | awk '/Function : .*k1_sym_pairsILb0ELb0E/{p=1} p{print} /Function : /{if(p&&!/k1_sym_pairsILb0ELb0E/)exit}' | grep -v '^\s*/\* 0x' > /tmp/_k1.sass
python3 - <<'PY'
import re
from collections import Counter
ins = []
for l in open('/tmp/_k1.sass'):
    m = re.search(r'/\*([0-9a-f]{4})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
for k, (addr, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)', t)
    if m and int(m.group(1), 16) < addr:
        body = [re.sub(r'^@!?U?P\d\s+', '', x) for a, x in ins if int(m.group(1), 16) <= a <= addr]
        c = Counter(x.split()[0].split('.')[0] for x in body)
        if len(body) > 150 and c.get('FFMA', 0) == 0 and c.get('FFMA2', 0) > 0:
            print(len(body), 'instructions:', dict(c.most_common(12)))
PY
echo "## first packed instructions of the planar loop (operand forms: .F32 = scalar broadcast, immediates folded)"
grep -E "FFMA2|FMUL2|FADD2" /tmp/_k1.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s+\/\*.*//' | awk 'NR>170 && NR<=200'
