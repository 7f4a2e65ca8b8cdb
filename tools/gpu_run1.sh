set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_run1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
# push curves
TRAIN_BUDGET_S=150 UTD=40 WINDOW=32768 RATE_WINDOW=25 SYNC_EVERY=8 timeout 400 python tools/train_curve.py push TD3_MLP 32 45 gpurun_out/r02_curve_push_td3_32_utd40.json > gpurun_out/r02_curve_push_A.log 2>&1
tail -3 gpurun_out/r02_curve_push_A.log | cut -c1-400
TRAIN_BUDGET_S=150 WINDOW=65536 timeout 400 python tools/train_curve.py push TD3_MLP 1024 700 gpurun_out/r02_curve_push_td3_1024_w64k.json > gpurun_out/r02_curve_push_B.log 2>&1
tail -3 gpurun_out/r02_curve_push_B.log | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 3 --quick --no-cpu > gpurun_out/r02_bench_k20.json 2> gpurun_out/r02_bench_k20.err
timeout 300 python bench.py --steps 2000 --warmup 20 --quick --no-cpu > gpurun_out/r02_bench_k2000.json 2> gpurun_out/r02_bench_k2000.err
cut -c1-300 gpurun_out/r02_bench_k20.json; cut -c1-300 gpurun_out/r02_bench_k2000.json
