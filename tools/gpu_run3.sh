set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_run3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run3_pytest.log
tail -8 gpurun_out/r02_run3_pytest.log
for k in 20 200 2000; do timeout 300 python bench.py --steps $k --warmup 3 --quick --no-cpu > gpurun_out/r02_bench_k$k.json 2> gpurun_out/r02_bench_k$k.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_k$k.json').read());print($k,d['ms_per_step'],d['timing'],d['e2e']['value'],d['config']['launch'])"; tail -3 gpurun_out/r02_bench_k$k.err; done
