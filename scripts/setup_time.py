import sys, time; sys.path.insert(0, '.')
import numpy as np, channelflow_b200 as cf
ctx = cf.Context(cf.GpuLib())
for (Nx,Ny,Nz) in ((512,257,512),(128,97,128)):
    nse = cf.Nse(ctx, Nx, Ny, Nz, 4*np.pi, 2*np.pi, -1.0, 1.0, Ubase=np.zeros(Ny), Wbase=np.zeros(Ny), nu=1/4000)
    for it in range(3):
        ctx.sync(); t0 = time.perf_counter(); nse.reset_lambda([11/6/0.002]); ctx.sync(); t1 = time.perf_counter()
        print((Nx,Ny,Nz), "tau setup ms", 1e3*(t1-t0))
