set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -x > gpurun_out/r02_run10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run10_pytest.log
tail -15 gpurun_out/r02_run10_pytest.log | cut -c1-300
python tools/policy_cost.py 4096 > gpurun_out/r02_policy_cost.txt 2>&1; tail -2 gpurun_out/r02_policy_cost.txt
