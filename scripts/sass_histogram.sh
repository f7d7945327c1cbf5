#!/bin/bash
# Opcode evidence for the Blackwell-native kernels: per kernel of libfz_fusion.so, how many tcgen05 / TMA / TMEM
# instructions its SASS holds (B200_PROFILING.md: UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA tensor
# load / store / reduce, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit).  Runs without a GPU.
SO=${1:-scikit-fusion_b200/libfz_fusion.so}
cuobjdump -sass "$SO" | awk '
  /Function : / { name=$3; sub(/^_ZN2fz[0-9]*/, "", name); fn=name }
  { n=split($0, f, /[ \t;]+/); for (i=1;i<=n;i++) if (f[i] ~ /^(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UTMAREDG|LDTM|STTM|UTCBAR|UTCATOMSWS|SYNCS|HMMA|DMMA|ELECT)/) { op=f[i]; sub(/\..*/, "", op); c[fn" "op]++ } }
  END { for (k in c) print k, c[k] }' | sort | awk '{ if ($1!=last) { if (last!="") print ""; printf "%s:", $1; last=$1 } printf " %s=%s", $2, $3 } END { print "" }' | grep -E "UTC|UTMA|LDTM"
