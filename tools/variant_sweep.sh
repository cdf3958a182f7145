#!/bin/bash
# Times arithmetic variants of the TM i8 kernel on every TM code (development aid).
# usage: variant_sweep.sh [variants...]   (default: 1 3 5 2)
V=${@:-1 3 5 2}
for spec in "8 65536 2.0" "7 65536 3.0" "6 65536 4.0" "5 262144 2.0" "4 262144 3.0"; do
  set -- $spec
  for a in $V; do
    LABRADOR_LDPC_TM_ARITH=$a python tools/quick_time.py $1 i8 $2 $3 | sed "s/^/arith $a: /" | cut -c1-140
  done
done
