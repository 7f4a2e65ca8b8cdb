set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_run12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run12_pytest.log
tail -4 gpurun_out/r02_run12_pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_run12_bench_k20.json 2> gpurun_out/r02_run12_bench_k20.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_run12_bench_k20.json').read().strip().splitlines()[-1]);print('k20',d['ms_per_step']*1e3, d['rollout_with_td3_actor']['us_per_rollout_step'])"
tail -3 gpurun_out/r02_run12_bench_k20.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:policy_mlp -s 6 -c 1 -f -o gpurun_out/r02_policy_v3 python tools/policy_cost.py 4096 > gpurun_out/r02_ncu_policy.log 2>&1; tail -2 gpurun_out/r02_ncu_policy.log
