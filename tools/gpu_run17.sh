set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:step_quad -s 300 -c 3 -f -o gpurun_out/r02_reach_n4096_quad python tools/profile_step.py reach 4096 1 304 > gpurun_out/r02_ncu_reach_quad.log 2>&1; tail -1 gpurun_out/r02_ncu_reach_quad.log
