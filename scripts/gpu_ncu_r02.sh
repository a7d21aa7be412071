# ncu evidence of round 2 (B200_PROFILING.md recipe: every command runs plain first, then under ncu).  The reports are
# summarised on the box (scripts/ncu_summary.py, scripts/refresh_profiles.py); only text travels back.
set -x
O=gpurun_out
# 1. launch list of the default bench command (sustained loop and comparator switched off: 32 launches of 4 kernels)
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --sustain-s 0 --no-comparator"
$CMD > $O/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_default.csv $CMD > $O/ncu_default.log 2>&1
# 2. the fused kernels, full set, 8 walkers (one launch each)
CMD="python bench.py --steps 1 --warmup 3 --walkers 8 --no-e2e --no-cpu --sustain-s 0 --no-comparator --no-graph"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv $CMD > $O/ncu1.log 2>&1
$CMD > $O/plain2.log 2>&1 && ncu --set full --clock-control none -k regex:k_fused_p -s 6 -c 2 -o $O/rep_fused $CMD > $O/ncu2.log 2>&1
python scripts/refresh_profiles.py $O/rep_fused.ncu-rep $O/launches.csv $O/launches_default.csv $O/plain_default.log r02 > $O/refresh.log 2>&1
mkdir -p $O/profiles && cp profiles/r02_ncu_fused_final_summary.txt profiles/r02_ncu_launches_*.csv profiles/r02_ncu_launches_default_bench_summary.txt profiles/ncu_traffic.json $O/profiles/
rm -f $O/rep_fused.ncu-rep
# 3. the generic two-pass kernels on config 2 (3 x 200^3 per walker, 4 walkers) and config 4 (16 x 512^3, 1 walker)
CMD="python bench.py --config 2 --walkers 4 --no-e2e --no-cpu --sustain-s 0 --no-comparator --steps 1 --no-graph"
$CMD > $O/plain3.log 2>&1 && ncu --set full --clock-control none -k regex:k_g2_p -s 6 -c 2 -o $O/rep_cfg2 $CMD > $O/ncu3.log 2>&1
(echo "ncu --set full --clock-control none -k regex:k_g2_p -s 6 -c 2  $CMD"; python scripts/ncu_summary.py $O/rep_cfg2.ncu-rep) > $O/profiles/r02_ncu_g2_cfg2_final_summary.txt 2>&1
rm -f $O/rep_cfg2.ncu-rep
CMD="python bench.py --config 4 --walkers 1 --no-e2e --no-cpu --sustain-s 0 --no-comparator --steps 1 --no-graph"
$CMD > $O/plain4.log 2>&1 && ncu --set full --clock-control none -k regex:k_g2_p -s 6 -c 2 -o $O/rep_cfg4 $CMD > $O/ncu4.log 2>&1
(echo "ncu --set full --clock-control none -k regex:k_g2_p -s 6 -c 2  $CMD"; python scripts/ncu_summary.py $O/rep_cfg4.ncu-rep) > $O/profiles/r02_ncu_g2_cfg4_final_summary.txt 2>&1
rm -f $O/rep_cfg4.ncu-rep
du -sh $O
