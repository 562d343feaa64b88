#!/bin/bash
# Blackwell-native evidence, tracked: per kernel of libemo_b200.so, how many tcgen05 (UTC*MMA), TMEM (LDTM / STTM),
# TMA (UTMALDG / UTMASTG), legacy tensor-core (HMMA) and cp.async (LDGSTS) instructions its SASS holds.
# Runs here (no GPU):  bash scripts/sass_evidence.sh > profiles/r02_sass_mnemonics.txt
set -u
SO=${1:-emo_disentanger_b200/libemo_b200.so}
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3 }
  /UTC[A-Z]*MMA/ { utc[fn]++ } /LDTM/ { ldtm[fn]++ } /STTM/ { sttm[fn]++ } /UTMALDG/ { tmal[fn]++ } /UTMASTG/ { tmas[fn]++ }
  /[ \t]HMMA/ { hmma[fn]++ } /LDGSTS/ { lds[fn]++ } { seen[fn]=1 }
  END { printf "%-8s %-6s %-6s %-8s %-8s %-6s %-7s %s\n", "UTCxMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "kernel";
        for (f in seen) if (utc[f] + tmal[f] + tmas[f] + hmma[f] + lds[f] > 0)
          printf "%-8d %-6d %-6d %-8d %-8d %-6d %-7d %s\n", utc[f], ldtm[f], sttm[f], tmal[f], tmas[f], hmma[f], lds[f], f }' | (read -r hdr; echo "$hdr"; sort -k8)
