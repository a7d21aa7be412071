#!/bin/bash
# Tuning sweep of the generic two-pass kernels (run on a GPU box): CTAs per SM in pass 1, x slabs per CTA; 1024-thread
# units on config 4.  Every variant twice, interleaved, to see the run-to-run spread of the box.
run() { python bench.py --config $2 --no-e2e --no-cpu --sustain-s 0 --no-comparator --no-graph --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), 'p1', round(k['fft_pass1'],3), 'p2', round(k['fft_last_pass_bin'],3))
"; }
for rep in 1 2; do
run "cfg2 default(minb4,slabs1)" 2
PS21_G2_P1_MINB3=1 run "cfg2 minb3,slabs1" 2
PS21_G2_SLABS=2 run "cfg2 minb4,slabs2" 2
PS21_G2_P1_MINB3=1 PS21_G2_SLABS=2 run "cfg2 minb3,slabs2" 2
done
for rep in 1 2; do
run "cfg4 default(512 threads)" 4
PS21_G2_P1_1024=1 run "cfg4 1024 threads radix<=8" 4
done
