set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:policy_mlp -s 6 -c 1 -f -o gpurun_out/r02_policy python tools/policy_cost.py 4096 > gpurun_out/r02_ncu_policy.log 2>&1; tail -2 gpurun_out/r02_ncu_policy.log
timeout 300 $NCU -k regex:step_lane -s 300 -c 3 -f -o gpurun_out/r02_reach_n4096 python tools/profile_step.py reach 4096 1 304 > gpurun_out/r02_ncu_reach.log 2>&1; tail -1 gpurun_out/r02_ncu_reach.log
timeout 300 $NCU -k regex:step_lane -s 200 -c 3 -f -o gpurun_out/r02_push_n4096 python tools/profile_step.py push 4096 1 204 > gpurun_out/r02_ncu_push.log 2>&1; tail -1 gpurun_out/r02_ncu_push.log
timeout 300 $NCU -k regex:step_lane -s 200 -c 3 -f -o gpurun_out/r02_pick_n2048 python tools/profile_step.py pick 2048 1 204 > gpurun_out/r02_ncu_pick.log 2>&1; tail -1 gpurun_out/r02_ncu_pick.log
ls -la gpurun_out/*.ncu-rep
