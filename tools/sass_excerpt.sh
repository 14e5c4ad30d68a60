#!/bin/bash
# SASS evidence for the claims DESIGN.md makes about the hot kernels (run here: cuobjdump needs no GPU):
#   tools/sass_excerpt.sh > profiles/sass_excerpts_r2.txt
# Per kernel: mnemonic histogram (top 16) and the first lines that show the instruction the claim is about.
cd "$(dirname "$0")/../pfilter-noetic_b200/csrc"
show() {   # kernel-name-substring  object  regex-of-interest  what
  echo "================================================================================"
  echo "$1   ($2)   -- $4"
  local body
  body=$(cuobjdump -sass "$2" 2>/dev/null | awk -v k="$1" '/Function :/ {on = index($0,k)>0} on')
  echo "-- mnemonic histogram (top 16)"
  echo "$body" | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -16 | awk '{printf "%s %s; ", $1, $2} END {print ""}'
  echo "-- first 8 instructions matching /$3/ (and their count)"
  echo "$body" | grep -E "$3" | head -8 | sed 's/^\s*//' | cut -c1-120
  echo "count: $(echo "$body" | grep -cE "$3")"
}
show k_sector_extract extract.o "FADD2|FFMA2|FMUL2|REDUX|LDGSTS" "packed fp32 pairs for the curvature taps, redux for the greedy pick, cp.async gather"
show k_ring_classify extract.o "MUFU.RSQ|VOTE|MATCH" "one MUFU.RSQ per point, tile purity from votes"
show k_normal_eq_stream solve.o "UBLKCP|SYNCS|DFMA|DADD|DMUL" "bulk (TMA) copies with mbarrier completion; fp64 without contraction"
show k_lm_solve solve.o "UCGABAR|MAPA|LD.E.*\[UR" "cluster barrier and distributed shared memory reads"
show k_mm_count merge.o "LDG.E.*128|EF|LTC" "16-byte loads with L2 eviction-priority hints"
show k_mm_write merge.o "LDG.E.*128|STG.E.*128|ACQBULK|GRIDDEP" "16-byte loads / stores, programmatic dependent launch"
show k_assoc_match match.o "REDUX|DFMA|MUFU" "redux merges of the per-lane top-5, fp64 fit in registers"
show k_sort_small primitives.o "MATCH|ATOMS" "match.any ranking, keys ping-pong in shared memory"
