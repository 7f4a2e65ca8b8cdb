set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
HANG_DUMP_S=100 ALGO=DARC_MLP timeout 150 $TR tools/dist_train_check.py > gpurun_out/r02_dist_check_8gpu_DARC_MLP.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dist_check_8gpu_DARC_MLP.log; grep "updates\|DIST_TRAIN_OK\|^rc=" gpurun_out/r02_dist_check_8gpu_DARC_MLP.log
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_8gpu_k20.json 2> gpurun_out/r02_bench_8gpu_k20.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_8gpu_k20.json').read().strip().splitlines()[-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'],json.dumps(d.get('train_update'))[:1500])"
tail -3 gpurun_out/r02_bench_8gpu_k20.err
TRAIN_BUDGET_S=30 timeout 200 $TR tools/train_curve.py reach DARC_MLP 4096 30 gpurun_out/r02_curve_reach_darc_8x4096.json > gpurun_out/r02_curve_darc8.log 2>&1; echo "curve rc=$?"; tail -2 gpurun_out/r02_curve_darc8.log | cut -c1-500
