# ncu evidence for the fused kernels (B200_PROFILING.md recipe): plain run first, then launch list, then full set
set -x
CMD="python bench.py --steps 1 --warmup 3 --walkers 8 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused_p -s 6 -c 2 -o gpurun_out/prof_fused_final5 $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
