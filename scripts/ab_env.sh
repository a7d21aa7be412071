#!/bin/bash
# A/B of one environment switch on the generic-kernel configurations, interleaved.  usage: ab_env.sh VAR=VALUE
run() { env $1 python bench.py --config $2 --no-e2e --no-cpu --sustain-s 0 --no-comparator --no-graph --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernel_ms_per_step']; print('[$1] cfg$2', round(d['ms_per_step'],3), 'p1', round(k['fft_pass1'],3), 'p2', round(k['fft_last_pass_bin'],3), 'lnl', d.get('self_check_lnl_walker0'))
"; }
for rep in 1 2; do for c in 2 4; do run "$1" $c; run "A=1" $c; done; done
for v in "$1" "A=1"; do env $v python scripts/time_shape.py 256 1000 4 2>&1 | tail -1 | cut -c1-330; done
