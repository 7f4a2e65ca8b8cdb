set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_run4_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run4_pytest.log
tail -8 gpurun_out/r02_run4_pytest.log
for k in 20 2000; do timeout 300 python bench.py --steps $k --warmup 3 --quick --no-cpu > gpurun_out/r02_bench_k$k.json 2> gpurun_out/r02_bench_k$k.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_k$k.json').read());print($k,d['ms_per_step'],d.get('timing'),d['e2e']['value'],d['config'].get('launch'))"; tail -3 gpurun_out/r02_bench_k$k.err; done
timeout 900 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
tail -c 2500 gpurun_out/r02_bench_full.json; tail -3 gpurun_out/r02_bench_full.err
TRAIN_BUDGET_S=150 WINDOW=65536 timeout 400 python tools/train_curve.py push TD3_MLP 1024 700 gpurun_out/r02_curve_push_td3_1024_w64k.json > gpurun_out/r02_curve_push_B.log 2>&1
tail -3 gpurun_out/r02_curve_push_B.log | cut -c1-600
TRAIN_BUDGET_S=150 WINDOW=16384 timeout 400 python tools/train_curve.py pick DATD3_MLP 2048 150 gpurun_out/r02_curve_pick_datd3_2048.json > gpurun_out/r02_curve_pick.log 2>&1
tail -2 gpurun_out/r02_curve_pick.log | cut -c1-600
