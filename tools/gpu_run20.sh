set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run20_pytest.log
tail -4 gpurun_out/r02_run20_pytest.log | cut -c1-300
for k in 20 2000; do timeout 300 python bench.py --steps $k --warmup 5 --quick --no-cpu > gpurun_out/r02_run20_bench_k$k.json 2> gpurun_out/r02_run20_bench_k$k.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_run20_bench_k$k.json').read().strip().splitlines()[-1]);print($k,d['ms_per_step']*1e3,'e2e us',4096/d['e2e']['value']*1e6)"; tail -2 gpurun_out/r02_run20_bench_k$k.err; done
for t in "push 4096" "pick 2048"; do set -- $t; timeout 200 python bench.py --task $1 --n-envs $2 --steps 500 --warmup 5 --quick --no-cpu > gpurun_out/r02_run20_$1.json 2> gpurun_out/r02_run20_$1.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_run20_$1.json').read().strip().splitlines()[-1]);print('$1',d['ms_per_step']*1e3,'us')"; done
