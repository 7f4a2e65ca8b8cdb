set -x
mkdir -p gpurun_out
TRAIN_BUDGET_S=170 WINDOW=16384 timeout 400 python tools/train_curve.py pick DATD3_MLP 2048 1200 gpurun_out/r02_curve_pick_datd3_2048.json > gpurun_out/r02_curve_pick.log 2>&1
tail -2 gpurun_out/r02_curve_pick.log | cut -c1-400
TRAIN_BUDGET_S=40 timeout 200 python tools/train_curve.py reach TD3_MLP 1024 100 gpurun_out/r02_curve_reach_td3_1024.json > gpurun_out/r02_curve_reach.log 2>&1
tail -2 gpurun_out/r02_curve_reach.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --cache-control none -s 4000 -c 300 --csv --log-file gpurun_out/r02_launches_bench_reach4096.csv python bench.py --steps 200 --warmup 3 --quick --no-cpu > gpurun_out/r02_ncu_launchlist.log 2>&1
tail -3 gpurun_out/r02_launches_bench_reach4096.csv | cut -c1-300
