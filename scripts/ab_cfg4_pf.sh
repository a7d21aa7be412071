#!/bin/bash
# where to issue the L2 prefetch of the next unit in the 512^3 pass 1 (PS21_PF_WHERE 0 top / 1 after z / 2 before the last
# y stage), with the reload in 2 or 3 pieces; every variant twice
run() { python bench.py --config 4 --no-e2e --no-cpu --sustain-s 0 --no-comparator --no-graph --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), 'p1', round(k['fft_pass1'],3), 'p2', round(k['fft_last_pass_bin'],3))
"; }
for rep in 1 2; do
PS21_L2_PREFETCH=0 run "off"
for w in 0 1 2; do
  PS21_PF_WHERE=$w run "where=$w"
done
PS21_PF_WHERE=2 PS21_G2_PARTS=3 run "where=2 parts=3"
PS21_PF_WHERE=1 PS21_G2_PARTS=3 run "where=1 parts=3"
done
