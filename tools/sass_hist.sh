#!/bin/bash
# usage: tools/sass_hist.sh <binary-or-.so> <mangled-name-substring>   -> SASS opcode histogram of the matching kernel(s)
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '
/Function : /{f=$3; on = (index(f, pat) > 0); if (on) print "## " f}
on && /^[ \t]+\/\*[0-9a-f]+\*\/[ \t]+/ { op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); c[f" "op]++; tot[f]++ }
END{for(k in c) print c[k], k; for (f in tot) print tot[f], f, "TOTAL"}' | sort -k2,2 -k1,1rn
