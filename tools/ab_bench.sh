#!/bin/bash
# A/B of an environment switch on the same box, interleaved (clock / power state drifts between runs):
#   tools/ab_bench.sh VAR [reps]  ->  gpurun_out/ab_<VAR>.txt with ms_per_step and SM clock of every run
VAR=$1; REPS=${2:-3}
OUT=gpurun_out/ab_${VAR}.txt; : > $OUT
for i in $(seq $REPS); do
  for v in 0 1; do
    if [ $v = 1 ]; then export $VAR=1; else unset $VAR; fi
    python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$VAR=$v', 'ms %.3f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], 'clk', d['clocks']['sm_mhz'], 'tail_ms %.4f' % d['roofline_tail']['ms'], 'conv %.3f' % d['extras']['kernel_ms_per_step']['conv'])" >> $OUT
  done
done
cat $OUT
