set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run2_pytest.log
tail -15 gpurun_out/r02_run2_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --quick --no-cpu > gpurun_out/r02_bench_k20.json 2> gpurun_out/r02_bench_k20.err
timeout 300 python bench.py --steps 2000 --warmup 20 --quick --no-cpu > gpurun_out/r02_bench_k2000.json 2> gpurun_out/r02_bench_k2000.err
cut -c1-300 gpurun_out/r02_bench_k20.json; cut -c1-300 gpurun_out/r02_bench_k2000.json
timeout 900 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
tail -c 3000 gpurun_out/r02_bench_full.json
# pick + DATD3 curve (BASELINE config 4), reach TD3 regression check
TRAIN_BUDGET_S=200 WINDOW=16384 timeout 500 python tools/train_curve.py pick DATD3_MLP 2048 150 gpurun_out/r02_curve_pick_datd3_2048.json > gpurun_out/r02_curve_pick.log 2>&1
tail -2 gpurun_out/r02_curve_pick.log | cut -c1-600
TRAIN_BUDGET_S=60 timeout 300 python tools/train_curve.py reach TD3_MLP 1024 100 gpurun_out/r02_curve_reach_td3_1024.json > gpurun_out/r02_curve_reach.log 2>&1
tail -2 gpurun_out/r02_curve_reach.log | cut -c1-600
