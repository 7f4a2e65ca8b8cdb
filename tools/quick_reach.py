import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import drl_on_robot_arm_b200 as pkg
dev=torch.device('cuda:0')
stream=torch.cuda.Stream()
for n,pool,k in ((4096,547,1000),(1<<20,3,60),(1<<22,1,20)):
    envs=[pkg.BatchedArmEnv('reach',n_envs=n,device=dev,seed=0,auto_reset=True,env_id_offset=b*n) for b in range(pool)]
    na=7 if n>=(1<<20) else 61
    a=(torch.rand((na,n,3),device=dev)*1.4-0.7)
    with torch.cuda.stream(stream):
        for j in range(8): envs[j%pool].step(a[j%na])
    stream.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g,stream=stream):
        for j in range(k): envs[j%pool].step(a[j%na])
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    best=1e9
    with torch.cuda.stream(stream):
        g.replay(); stream.synchronize()
        for r in range(3):
            e0.record(stream); g.replay(); e1.record(stream); stream.synchronize()
            best=min(best,e0.elapsed_time(e1)*1e3/k)
    print(n, 'us/launch %.3f'%best, 'G env-steps/s %.3f'%(n/best/1e3), flush=True)
    del g
    for e in envs: e.close()
