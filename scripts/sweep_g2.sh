#!/bin/bash
# Tuning sweep of the generic two-pass kernels on BASELINE config 2 (run on a GPU box): CTAs per SM in pass 1, x slabs
# per CTA, tiles per pass-2 CTA.
run() { python bench.py --config 2 --no-e2e --no-cpu --sustain-s 0 --no-comparator --no-graph --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']; print('$1', round(d['ms_per_step'],3), 'p1', round(k['fft_pass1'],3), 'p2', round(k['fft_last_pass_bin'],3))
"; }
run base
PS21_G2_P1_MINB4=1 run minb4
PS21_G2_SLABS=1 run slabs1
PS21_G2_SLABS=4 run slabs4
PS21_G2_SLABS=8 run slabs8
PS21_G2_TILES_PER_CTA=6 run per6
PS21_G2_TILES_PER_CTA=24 run per24
PS21_G2_TILES_PER_CTA=50 run per50
