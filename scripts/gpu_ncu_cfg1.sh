# ncu capture of the 128^2-plane fused kernels (config 1), plain run first
set -x
CMD="python bench.py --config 1 --steps 1 --warmup 3 --walkers 16 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_c1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fused_p -s 6 -c 2 -o gpurun_out/prof_cfg1 $CMD > gpurun_out/ncu_c1.log 2>&1
tail -2 gpurun_out/ncu_c1.log
