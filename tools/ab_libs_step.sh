#!/bin/bash
# Same-box A/B of two builds (ab_libs/lib_prev.so vs ab_libs/lib_new.so): graph-replayed step time, interleaved processes.
REPS=${1:-3}
for i in $(seq $REPS); do
  for v in prev new; do
    cp ab_libs/lib_$v.so mb_istft_vits_b200/libmbistft.so
    echo -n "$v "; python tools/ab_flags.py --a 0 --b 0 --blocks 3 2>/dev/null | grep median | head -1
  done
done
cp ab_libs/lib_new.so mb_istft_vits_b200/libmbistft.so
