#!/bin/bash
# Run tools/time_one.py for every variant library given on the command line (names under 1-stage-wseg_b200/variants/).
# usage: tools/sweep_variants.sh OUT.log H W B name1 name2 ...
out=$1; H=$2; W=$3; B=$4; shift 4
: > "$out"
for v in "$@"; do
  echo "== $v" >> "$out"
  PAMR_B200_LIB=$PWD/1-stage-wseg_b200/variants/$v.so timeout 120 python tools/time_one.py $H $W $B 2>&1 | grep -E "marginal|Error|error" >> "$out"
done
cat "$out"
