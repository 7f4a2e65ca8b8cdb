set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_run21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run21_pytest.log
tail -12 gpurun_out/r02_run21_pytest.log | cut -c1-300
for t in "push 1048576" "pick 1048576" "push 4096" "pick 2048"; do set -- $t; timeout 200 python bench.py --task $1 --n-envs $2 --steps 40 --warmup 5 --quick --no-cpu > gpurun_out/r02_run21_$1_$2.json 2> gpurun_out/r02_run21_$1.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_run21_$1_$2.json').read().strip().splitlines()[-1]);print('$1 $2',d['ms_per_step']*1e3,'us')"; tail -1 gpurun_out/r02_run21_$1.err; done
