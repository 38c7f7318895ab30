#!/bin/bash
# tools/sass_opcodes.sh > profiles/sass_opcodes.txt -- counts of the Blackwell-specific opcodes per kernel of the shipped
# library (cuobjdump -sass sprl_b200/lib/libsprl_b200.so): tcgen05 MMAs (UTCHMMA), TMEM loads / stores (LDTM / STTM),
# bulk copies (UBLKCP), tensor-core barriers (UTCBAR), mbarrier ops (SYNCS), warp reductions (REDUX).
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass sprl_b200/lib/libsprl_b200.so ($(date -u +%Y-%m-%dT%H:%MZ), $(git rev-parse --short HEAD))"
cuobjdump -sass sprl_b200/lib/libsprl_b200.so | awk '
/Function :/ { fn=$3 }
/UTCHMMA|UTCQMMA|UTCMMA/ { c[fn,"UTCHMMA"]++ }
/LDTM/ { c[fn,"LDTM"]++ }
/STTM/ { c[fn,"STTM"]++ }
/UBLKCP|UBLKPF/ { c[fn,"UBLKCP"]++ }
/UTCBAR/ { c[fn,"UTCBAR"]++ }
/SYNCS/ { c[fn,"SYNCS"]++ }
/REDUX/ { c[fn,"REDUX"]++ }
/UTCATOMSWS|UTCCP/ { c[fn,"UTC_OTHER"]++ }
{ if (fn != "") n[fn]++ }
END {
  printf "%-10s %-8s %-6s %-6s %-7s %-7s %-6s %-6s %s\n", "sass_lines", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "REDUX", "kernel"
  for (f in n) printf "%-10d %-8d %-6d %-6d %-7d %-7d %-6d %-6d %s\n", n[f], c[f,"UTCHMMA"], c[f,"LDTM"], c[f,"STTM"], c[f,"UBLKCP"], c[f,"UTCBAR"], c[f,"SYNCS"], c[f,"REDUX"], f
}' | (read -r hdr; echo "$hdr"; sort -k9)
