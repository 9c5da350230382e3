#!/bin/bash
# Regenerates profiles/r1_sass_evidence.txt: opcode counts that show what the kernels are built from (UBLKCP = TMA bulk
# copy, SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed FP32x2, REDG = the integer accumulator atomics, MEMBAR.*.SYS =
# the system-scope fences of the peer barrier).  Run from the repo root after __graft_entry__.build().
LIB=carla-social-force-model_b200/sfm_b200/libsfm_b200.so
for k in k1_sym_pairsILb0 k2_segmentsILi0 k3_integrate k7_barrier k1_sym_finishILb0; do
  echo "## $k"
  cuobjdump -sass $LIB | awk -v k="$k" '/Function :/{f=($0 ~ k)} f' | grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" |
    awk '{print $NF}' | sed -E 's/\..*//' | sort | uniq -c | sort -rn | head -16 | tr '\n' ';'
  echo
done
