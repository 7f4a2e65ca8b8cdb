set -x
mkdir -p gpurun_out
python tools/policy_cost.py 4096 2>&1 | tail -2
for lib in libarmsim.so libarmsim_oldcube.so; do for t in "push 4096" "pick 2048"; do set -- $t; ARMSIM_LIB=$PWD/drl-on-robot-arm_b200/$lib timeout 200 python bench.py --task $1 --n-envs $2 --steps 500 --warmup 5 --quick --no-cpu > gpurun_out/r02_run8_$1_$lib.json 2> gpurun_out/r02_run8_$1.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_run8_$1_$lib.json').read().strip().splitlines()[-1]);print('$lib $1',d['ms_per_step']*1e3,'us', d['timing']['p10_ms_per_step']*1e3, d['timing']['p90_ms_per_step']*1e3)"; done; done
