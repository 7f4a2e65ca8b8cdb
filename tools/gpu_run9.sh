set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_run9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run9_pytest.log
tail -4 gpurun_out/r02_run9_pytest.log | cut -c1-300
python tools/policy_cost.py 4096 > gpurun_out/r02_policy_cost.txt 2>&1; tail -2 gpurun_out/r02_policy_cost.txt
for t in "push 4096" "pick 2048"; do set -- $t; timeout 200 python bench.py --task $1 --n-envs $2 --steps 500 --warmup 5 --quick --no-cpu > gpurun_out/r02_run9_$1.json 2> gpurun_out/r02_run9_$1.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_run9_$1.json').read().strip().splitlines()[-1]);print('$1',d['ms_per_step']*1e3,'us', d['timing']['p10_ms_per_step']*1e3, d['timing']['p90_ms_per_step']*1e3)"; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_run9_bench_k20.json 2> gpurun_out/r02_run9_bench_k20.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_run9_bench_k20.json').read().strip().splitlines()[-1]);print('k20',d['ms_per_step']*1e3, d['rollout_with_td3_actor']['us_per_rollout_step']); print({k:(v['us_per_launch']) for k,v in d['other_configs'].items()})"
tail -3 gpurun_out/r02_run9_bench_k20.err
