# Round-end evidence pass on one B200 (gpurun -- 'bash tools/gpu_evidence.sh'): GPU tests, the bench lines, ncu captures.
# Outputs land in gpurun_out/; copy what should be judged into profiles/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log
tail -4 gpurun_out/r02_final_pytest.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 300 python bench.py --steps 20 --warmup 5 --quick --no-cpu > gpurun_out/r02_bench_1gpu_k20.json 2> gpurun_out/r02_bench_1gpu_k20.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
python -c "
import json
for f in ('r02_bench_1gpu','r02_bench_1gpu_k20','r02_bench_reference_arm'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'))"
NCU="ncu --set full --clock-control none --import-source on"
for t in "reach 4096 304 300" "push 4096 204 200" "pick 2048 204 200"; do set -- $t
  timeout 300 $NCU -k regex:step_lane -s $4 -c 3 -f -o gpurun_out/r02_$1_n$2 python tools/profile_step.py $1 $2 1 $3 > gpurun_out/r02_ncu_$1.log 2>&1; tail -1 gpurun_out/r02_ncu_$1.log
done
timeout 300 $NCU -k regex:policy_mlp -s 6 -c 1 -f -o gpurun_out/r02_policy python tools/policy_cost.py 4096 > gpurun_out/r02_ncu_policy.log 2>&1
python tools/policy_cost.py 4096 > gpurun_out/r02_policy_cost.txt 2>&1; tail -2 gpurun_out/r02_policy_cost.txt
