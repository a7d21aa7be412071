#!/bin/bash
# A/B of the L2 prefetch in the two-pass kernels (PS21_L2_PREFETCH = 0 off, 1 next unit / tile of the CTA, 3 also the
# first unit / tile of the succeeding CTA) on one GPU box: every variant twice, interleaved.
run() { python bench.py $2 --no-e2e --no-cpu --sustain-s 0 --no-comparator --no-graph --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']; print('$1', 'pf=${PS21_L2_PREFETCH:-0}', 'ahead=${PS21_PF_AHEAD:-auto}', round(d['ms_per_step'],3), 'p1', round(k['fft_pass1'],3), 'p2', round(k['fft_last_pass_bin'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
"; }
for rep in 1 2; do
for cfg in "" "--config 2" "--config 4" "--config 1"; do
  for pf in 0 1 3; do
    PS21_L2_PREFETCH=$pf run "cfg[${cfg:-default}]" "$cfg"
  done
done
done
