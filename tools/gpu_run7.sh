set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_run7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run7_pytest.log
tail -12 gpurun_out/r02_run7_pytest.log | cut -c1-300
for t in "reach 4096" "push 4096" "pick 2048"; do set -- $t; timeout 200 python bench.py --task $1 --n-envs $2 --steps 200 --warmup 5 --quick --no-cpu > gpurun_out/r02_run7_$1.json 2> gpurun_out/r02_run7_$1.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_run7_$1.json').read().strip().splitlines()[-1]);print('$1',d['ms_per_step']*1e3,'us', d['timing']['p10_ms_per_step']*1e3, d['timing']['p90_ms_per_step']*1e3, 'e2e us', $2/d['e2e']['value']*1e6)"; done
for mg in 148 296; do ARMSIM_MIN_GRID=$mg timeout 200 python bench.py --task reach --n-envs 4096 --steps 200 --warmup 5 --quick --no-cpu > gpurun_out/r02_run7_reach_mg$mg.json 2> gpurun_out/r02_run7_reach_mg$mg.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_run7_reach_mg$mg.json').read().strip().splitlines()[-1]);print('reach min_grid $mg',d['ms_per_step']*1e3,'us', 'e2e us', 4096/d['e2e']['value']*1e6)"; done
ARMSIM_MIN_GRID=148 timeout 200 python bench.py --task push --n-envs 4096 --steps 200 --warmup 5 --quick --no-cpu > gpurun_out/r02_run7_push_mg148.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r02_run7_push_mg148.json').read().strip().splitlines()[-1]);print('push min_grid 148',d['ms_per_step']*1e3,'us')"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_run7_bench_k20.json 2> gpurun_out/r02_run7_bench_k20.err
python -c "
import json;d=json.loads(open('gpurun_out/r02_run7_bench_k20.json').read().strip().splitlines()[-1]);print('k20',d['ms_per_step']*1e3, d['rollout_with_td3_actor']); print({k:(v['us_per_launch']) for k,v in d['other_configs'].items()})"
