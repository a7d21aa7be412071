# launch list of the default bench command (B200_PROFILING.md: plain run first, same command line under ncu)
set -x
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_default.log 2>&1
tail -2 gpurun_out/plain_default.log | cut -c1-300
