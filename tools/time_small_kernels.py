"""Times the small field kernels around the sweeps (MPCFL, total mass, midpoint, BCf!) at 512^3 Float32 on cuda:0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import interfaceadvection.jl_b200 as ia

N = (512, 512, 512)
Ng = tuple(n + 2 for n in N)
T = torch.float32
f = ia.jl_zeros(Ng, T, "cuda"); f.uniform_()
u = ia.jl_zeros(Ng + (3,), T, "cuda"); u.uniform_(-0.3, 0.3)
ctx = ia.context_for(f)
s = torch.cuda.current_stream().cuda_stream


def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print("MPCFL          %.3f ms (reads 1.63 GB)" % t(lambda: ctx.mpcfl(s, u.data_ptr())))
print("sum_inside     %.3f ms (reads 0.54 GB)" % t(lambda: ia.sum_inside(f)))
g = ia.jl_zeros(Ng, T, "cuda")
print("axpby midpoint %.3f ms (1.63 GB)" % t(lambda: ctx.axpby(s, g.data_ptr(), 0.5, g.data_ptr(), 0.5, f.data_ptr())))
print("BCf            %.3f ms" % t(lambda: ia.BCf(f, (1, 2))))
