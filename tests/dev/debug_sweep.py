"""Debug helper (GPU): compare the CUDA CMOM / VOF sweeps with the oracle and print where they differ."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import pyoracle as O
from tests.helpers import make_state, alloc_cmom, oracle_cmom_call, inside
from tests.test_gpu_parity import run_cuda_cmom
import interfaceadvection.jl_b200 as ia

def where(err, tol, name):
    idx = np.argwhere(err > tol)
    if len(idx) == 0:
        print(f"   {name}: OK (max {err.max():.3e})"); return
    lo, hi = idx.min(0), idx.max(0)
    print(f"   {name}: {len(idx)} bad cells of {err.size}, max {err.max():.3e}, index range lo={lo+1} hi={hi+1} (1-based incl ghost), first={idx[0]+1}")
    for ax in range(err.ndim):
        cnt = np.bincount(idx[:, ax], minlength=err.shape[ax])
        print(f"      axis{ax}: " + " ".join(str(c) for c in cnt))

def case(N, kind, perdir, T, dirO, tol):
    st = make_state(N, kind, T, perdir=perdir, uBC=(0.0,) * len(N))
    a0 = alloc_cmom(st)
    O.u2rhou(a0["rhou"], st["u"], st["f"], st["lam_rho"]); O.BC(a0["rhou"], st["uBC"], False, st["perdir"])
    rhou0 = a0["rhou"].copy(order="F")
    f_o = st["f"].copy(order="F")
    so, rep, ao = oracle_cmom_call(st, f_o, st["u"], st["u"], st["u"], rhou0, 1.0, dirO)
    sc, f_c, ru_c = run_cuda_cmom(ia, st, st["f"], st["u"], st["u"], st["u"], rhou0, 1.0, dirO)
    print(f"case N={N} kind={kind} per={perdir} T={T.__name__} dirO={dirO} status cuda={sc} oracle={so}")
    where(np.abs(f_c - f_o), tol, "f")
    for d in range(len(N)):
        where(np.abs(ru_c[..., d] - ao["rhou"][..., d])[tuple([slice(1, -1)] * len(N))], tol, f"rhou{d+1}(inside)")

if __name__ == "__main__":
    T = np.float64
    for N, kind, per in [((16, 12, 10), "C2", ()), ((16, 12, 10), "C3", (1, 2, 3)), ((40, 24, 44), "C3", ())]:
        for dirO in [(3, 1, 2), (2, 3, 1)]:
            case(N, kind, per, T, dirO, 1e-12)
