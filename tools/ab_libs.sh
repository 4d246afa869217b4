#!/bin/bash
# Same-box A/B of two builds of libmbistft.so (ab_libs/lib_prev.so vs ab_libs/lib_new.so), interleaved.
REPS=${1:-3}
OUT=gpurun_out/ab_libs.txt; : > $OUT
for i in $(seq $REPS); do
  for v in prev new; do
    cp ab_libs/lib_$v.so mb_istft_vits_b200/libmbistft.so
    python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$v', 'ms %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'clk', d['clocks']['sm_mhz'], 'conv %.3f' % d['extras']['kernel_ms_per_step']['conv'])" >> $OUT
  done
done
cp ab_libs/lib_new.so mb_istft_vits_b200/libmbistft.so
cat $OUT
