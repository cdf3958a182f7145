#!/bin/bash
for spec in "3 i8 262144 4.0" "3 f32 65536 4.0" "4 i16 65536 3.0" "4 f32 65536 3.0" "5 i16 65536 2.0" "5 i32 65536 2.0" "5 f32 65536 2.0" "6 i16 16384 4.0" "6 f32 16384 4.0" "7 i16 16384 3.0" "7 f32 16384 3.0" "8 i16 8192 2.0" "8 f32 8192 2.0"; do
  set -- $spec
  python tools/quick_time.py $1 $2 $3 $4 | cut -c1-120
  LABRADOR_LDPC_FORCE_GENERIC=1 python tools/quick_time.py $1 $2 $3 $4 | cut -c1-120
done
