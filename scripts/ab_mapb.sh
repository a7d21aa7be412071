#!/bin/bash
# A/B of the lane mapping of the last x stage in the generic pass 2 (PS21_G2_NOMAPB=1: lanes = (k_z, ky_b) as before)
run() { python bench.py --config $2 --no-e2e --no-cpu --sustain-s 0 --no-comparator --no-graph --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), 'p1', round(k['fft_pass1'],3), 'p2', round(k['fft_last_pass_bin'],3), 'lnl', d.get('self_check_lnl_walker0'))
"; }
for rep in 1 2; do
for c in 2 4; do
PS21_G2_NOMAPB=1 run "cfg$c old-map" $c
run "cfg$c new-map" $c
done
done
for v in "PS21_G2_NOMAPB=1" "A=1"; do for s in "256 1000 4" "256 900 4" "128 1237 5"; do echo "[$v]"; env $v python scripts/time_shape.py $s 2>&1 | tail -1 | cut -c1-400; done; done
