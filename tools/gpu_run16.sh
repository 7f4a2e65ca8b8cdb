set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run16_pytest.log
tail -12 gpurun_out/r02_run16_pytest.log | cut -c1-300
for k in 20 2000; do timeout 300 python bench.py --steps $k --warmup 5 --quick --no-cpu > gpurun_out/r02_run16_bench_k$k.json 2> gpurun_out/r02_run16_bench_k$k.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_run16_bench_k$k.json').read().strip().splitlines()[-1]);print($k,d['ms_per_step']*1e3,d['config']['mapping'],'e2e us',4096/d['e2e']['value']*1e6)"; tail -2 gpurun_out/r02_run16_bench_k$k.err; done
