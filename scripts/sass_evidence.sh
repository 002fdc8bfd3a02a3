#!/bin/bash
# Blackwell-specific SASS mnemonics per kernel of the in-tree extension (run where cuobjdump is):
#   bash scripts/sass_evidence.sh > profiles/sass/blackwell_mnemonics_by_kernel.txt
SO=${1:-torchdistpackage_b200/_C.so}
cuobjdump -sass "$SO" 2>/dev/null | awk '
  /Function :/ { fn=$3; sub(/^_ZN3tdp/,"",fn); fn=substr(fn,1,70) }
  /UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UTMACCTL|LDTM|STTM|UTCBAR|UTCATOMSWS|UBLKCP|\.MC|LDGMC|MULTIMEM|FFMA2|FMUL2|FADD2|REDG.*SYS|ATOMG.*SYS|STG.*STRONG\.SYS|LDG.*STRONG\.SYS|MUFU\.TANH|SYNCS|UCGABAR|CGAERRBAR|ACQBULK/ {
    line=$0; sub(/^[ \t]*\/\*[0-9a-f]+\*\/[ \t]*/,"",line); sub(/;.*/,"",line);
    n=split(line,t," "); op=t[1]; if (op ~ /^@/) op=t[2];
    cnt[fn" : "op]++ }
  END { for (k in cnt) printf "%6d  %s\n", cnt[k], k }' | sort -k2
