#!/bin/bash
# SASS evidence for the Blackwell-native kernels: per kernel, the count of tcgen05 (UTC*MMA) / TMEM (LDTM) / TMA
# (UTMALDG, UTMASTG) / multimem (LDGMC) / system-scope peer atomics + vector reductions
# (B200_PROFILING.md "What proves a Blackwell-native kernel").
LIB=${1:-dynamic_load_balance_distributeddnn_b200/libdlb_b200.so}
cuobjdump -sass "$LIB" | awk '
/Function :/ {fn=$3}
{ m=$2; if (m ~ /^@/) m=$3 }   # predicated instructions carry the guard in column 2
m ~ /^(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|UTCATOMSWS|LDGMC|REDG|ATOMG\.E\.CAS\.STRONG\.SYS|STG\.E\.128\.STRONG\.SYS|LDG\.E\.128\.STRONG\.SYS|SYNCS\.ARRIVE|SYNCS\.PHASECHK)/ { c[fn" "m]++ }
END { for (k in c) { split(k, a, " "); printf "%s %s %d\n", a[1], a[2], c[k] } }' | sort | while read fn m n; do printf "%-6d %-42s %s\n" "$n" "$m" "$(echo $fn | c++filt | sed -E 's/\(anonymous namespace\):://g; s/\(CUtensorMap_st.*//; s/\(CommArgs.*//')"; done
