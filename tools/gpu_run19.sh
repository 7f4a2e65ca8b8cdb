set -x
mkdir -p gpurun_out
TRAIN_BUDGET_S=60 UTD=40 WINDOW=32768 RATE_WINDOW=25 SYNC_EVERY=8 timeout 200 python tools/train_curve.py reach TD3_MLP 16 40 gpurun_out/r02_curve_reach_td3_16_utd40.json > gpurun_out/r02_curve_reach_utd.log 2>&1
tail -1 gpurun_out/r02_curve_reach_utd.log | cut -c1-900
TRAIN_BUDGET_S=100 UTD=40 WINDOW=65536 RATE_WINDOW=25 SYNC_EVERY=8 timeout 300 python tools/train_curve.py push TD3_MLP 16 100 gpurun_out/r02_curve_push_td3_16_utd40.json > gpurun_out/r02_curve_push_utd.log 2>&1
tail -1 gpurun_out/r02_curve_push_utd.log | cut -c1-1500
