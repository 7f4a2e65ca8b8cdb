set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_run18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run18_pytest.log
tail -6 gpurun_out/r02_run18_pytest.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run18_bench_k20.json 2> gpurun_out/r02_run18_bench_k20.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_run18_bench_k20.json').read().strip().splitlines()[-1]);print('k20',d['ms_per_step']*1e3, d['e2e']['value'], d['rollout_with_td3_actor']['us_per_rollout_step'], d['cpu_baseline']['value'])"
tail -3 gpurun_out/r02_run18_bench_k20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_run18_bench_ref.json 2> gpurun_out/r02_run18_bench_ref.err; cut -c1-600 gpurun_out/r02_run18_bench_ref.json
TRAIN_BUDGET_S=30 timeout 200 python tools/train_curve.py reach TD3_MLP 1024 100 gpurun_out/r02_curve_reach_td3_1024.json > gpurun_out/r02_curve_reach.log 2>&1
tail -1 gpurun_out/r02_curve_reach.log | cut -c1-300
