import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'mst-cfd_b200')
import numpy as np, argparse
import bench, mstgpu
a = argparse.Namespace(workload='box', n=int(sys.argv[1]))
f,Q0,_ = bench.build_workload(a)
ctx = mstgpu.Context(f, order=2, flux='roe')
for dt in [1e-4, 5e-5, 2.5e-5, 1e-5]:
    ctx.set_state(Q0)
    out=[]
    for it in range(4):
        ctx.step(dt, 25)
        try:
            r = ctx.residual(); out.append('%.2e'%r[0])
        except Exception as e:
            out.append('NaN'); break
    print('dt',dt,'rho residual after 25,50,75,100 steps:',out, flush=True)
