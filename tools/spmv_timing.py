import sys, numpy as np
sys.path.insert(0,'/root/repo')
from navier_stokes_solver_b200 import binding as B
d=B.Disc.generate(300,100); dev=B.Device(d)
dev.upload(B.VEC_TMP0, np.random.default_rng(42).uniform(-1,1,d.n))
dev.set_time_params(B.MODE_NEWTON, 1/90.); dev.upload(B.VEC_SOLUTION, B.synthetic_state(d,1234)); dev.assemble(B.MODE_NEWTON, False, 1/90.)
for mode in (3,2):
    dev.set_option(B.OPT_STREAM_SPMV, mode)
    for w in (0,1):
        dev.time_kernel(w,5,True)
        print("mode",mode,"kernel",w,"flushed %.4f ms"%dev.time_kernel(w,50,1),"b2b %.4f ms"%dev.time_kernel(w,200,2), flush=True)
