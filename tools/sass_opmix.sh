#!/bin/bash
# SASS op mix per kernel: tools/sass_opmix.sh <pattern> [lib]   (counts static instructions by mnemonic)
LIB=${2:-idocp_b200/libidocp_b200.so}
cuobjdump -sass "$LIB" 2>/dev/null | awk -v pat="$1" '
/Function :/ {fn=$3; on = (fn ~ pat)}
on && /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+/ { op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); sub(/;$/,"",op); c[fn" "op]++; tot[fn]++ }
END { for (k in c) print c[k], k; for (f in tot) print tot[f], f, "TOTAL" }' | sort -k2,2 -k1,1nr
