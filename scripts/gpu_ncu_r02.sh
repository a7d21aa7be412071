# ncu evidence of round 2 (B200_PROFILING.md recipe: every command runs plain first, then under ncu)
set -x
O=gpurun_out
# 1. launch list of the default bench command (graph replay and the sustained loop switched off: 32 launches)
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --sustain-s 0 --no-comparator"
$CMD > $O/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv $CMD > $O/ncu_default.log 2>&1
# 2. the fused kernels, full set, 8 walkers (one launch each)
CMD="python bench.py --steps 1 --warmup 3 --walkers 8 --no-e2e --no-cpu --sustain-s 0 --no-comparator --no-graph"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv $CMD > $O/ncu1.log 2>&1
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused_p -s 6 -c 2 -o $O/r02_ncu_fused_final $CMD > $O/ncu2.log 2>&1
# 3. the generic two-pass kernels on config 2 (3 x 200^3 per walker, 4 walkers) and config 4 (16 x 512^3, 1 walker)
CMD="python bench.py --config 2 --walkers 4 --no-e2e --no-cpu --sustain-s 0 --no-comparator --steps 1 --no-graph"
$CMD > $O/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_g2_p -s 6 -c 2 -o $O/r02_ncu_g2_cfg2_final $CMD > $O/ncu3.log 2>&1
CMD="python bench.py --config 4 --walkers 1 --no-e2e --no-cpu --sustain-s 0 --no-comparator --steps 1 --no-graph"
$CMD > $O/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_g2_p -s 6 -c 2 -o $O/r02_ncu_g2_cfg4_final $CMD > $O/ncu4.log 2>&1
tail -2 $O/ncu4.log
