# ncu capture of the general-path kernels (plain run first)
set -x
CMD="python bench.py --steps 1 --warmup 3 --walkers 4 --path general --no-e2e --no-cpu"
$CMD > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_general_ -s 9 -c 3 -o gpurun_out/prof_general $CMD > gpurun_out/ncu_g.log 2>&1
tail -2 gpurun_out/ncu_g.log
