#!/bin/bash
# Histogram of the SASS mnemonics that show what the hot kernels are made of (TMA bulk copies = UBLKCP, mbarrier = SYNCS, named
# barriers = BAR, FP64 FMA = DFMA, shuffles, shared-memory loads), per kernel, from the shipped libnsx.so.
# usage: tools/sass_excerpt.sh > profiles/r02_sass_excerpt.md
LIB=${1:-navier_stokes_solver_b200/libnsx.so}
echo "# SASS of the shipped \`$(basename $LIB)\` (sm_100a): mnemonic counts per hot kernel"
echo
echo "\`cuobjdump -sass\` of the library the tests and bench.py load; UBLKCP = cp.async.bulk (1-D TMA copy), SYNCS = mbarrier operations,"
echo "BAR = bar.sync (CTA / named barriers), DFMA / DADD / DMUL = FP64 pipe, SHFL = warp shuffles, LDS = shared-memory loads."
echo
for k in k_spmv_tma k_sweep_block k_sweep_phased k_multi_dot2 k_multi_axpy_norm2 k_assemble k_fg_step; do
  echo "## $k"
  echo
  echo '```'
  cuobjdump -sass $LIB 2>/dev/null | awk -v k="$k" '/Function : /{f=index($0,k)>0; if(f) n++} f' | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" | awk '{print $NF}' | sed -E 's/^(UBLKCP|SYNCS|BAR|DFMA|DADD|DMUL|SHFL|LDS|STS|LDG|STG|ATOM|RED|MEMBAR|FENCE|UTMA[A-Z]*)[A-Z0-9_.]*/\1/' | grep -E "^(UBLKCP|SYNCS|BAR|DFMA|DADD|DMUL|SHFL|LDS|STS|LDG|STG|ATOM|RED|MEMBAR|FENCE|UTMA)" | sort | uniq -c | sort -rn | awk '{printf "%-8s %s\n", $2, $1}'
  echo '```'
  cuobjdump -sass $LIB 2>/dev/null | awk -v k="$k" '/Function : /{f=index($0,k)>0} f' | grep -E "UBLKCP|SYNCS" | head -4 | sed 's/^\s*/    /'
  echo
done
