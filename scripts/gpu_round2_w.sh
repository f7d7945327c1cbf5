#!/bin/bash
# GPU-box script (round 2, call W): shared-memory geometry study of the single-term fused kernel -- row blocks per group x relation
# stages in flight x flush staging (csrc/umma_fused1.cuh: FZ_F1_*), timing of each compiled variant, correctness of the candidates.
OUT=gpurun_out/${1:-r2w}
mkdir -p $OUT
D=scikit-fusion_b200/csrc/dev
for V in base b4s3h b3s4h b2s4h base; do
  echo "== $V" >> $OUT/geometry.log
  timeout 120 $D/umma_probe_$V b ${2:-36864} 3 148 >> $OUT/geometry.log 2>&1; echo "exit $?" >> $OUT/geometry.log
done
cat $OUT/geometry.log
for V in b3s4h b4s3h; do
  timeout 400 $D/umma_probe_$V 1 0 0 148 > $OUT/correct_$V.log 2>&1; echo "exit $?" >> $OUT/correct_$V.log
  grep -c OK $OUT/correct_$V.log; grep "FAIL\|failing\|exit" $OUT/correct_$V.log | head -5
done
