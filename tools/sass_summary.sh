#!/bin/bash
# SASS mnemonic counts per object of the built library (cuobjdump -sass): the instructions that prove which hardware paths
# a kernel file uses (UTCIMMA / LDTM / UTMALDG = tcgen05.mma, tcgen05.ld, TMA; LDGSTS = cp.async; IDP = dp4a; REDUX ...).
# usage: tools/sass_summary.sh > profiles/rNN_sass_mnemonics.csv     (after `make -C gravity2_b200/csrc`)
cd "$(dirname "$0")/../gravity2_b200/csrc" || exit 1
M="UTCIMMA UTCHMMA LDTM UTMALDG UTCBAR SYNCS LDGSTS IDP VIMNMX IMAD DFMA DMUL POPC REDUX PRMT MUFU.RSQ LDS.128 LDG.E STG.E BAR.SYNC"
echo "# cuobjdump -sass <object> | grep -c <mnemonic>; objects of gravity2_b200/csrc built with -gencode arch=compute_100a,code=sm_100a"
printf "object"; for m in $M; do printf ",%s" "$m"; done; echo
for o in *.o; do
    cuobjdump -sass "$o" > /tmp/sass_summary.txt 2>/dev/null
    printf "%s" "$o"
    for m in $M; do printf ",%s" "$(grep -c -- "$m" /tmp/sass_summary.txt)"; done
    echo
done
