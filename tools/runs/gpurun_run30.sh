O=gpurun_out
( time python -m pytest tests/test_gpu_post.py -q -x --durations=3 ) > $O/r2_s30_pytest.log 2>&1; tail -25 $O/r2_s30_pytest.log | cut -c1-300
python - > $O/r2_s30_post_timing.json <<'PY'
import json, torch, numpy as np, sys
sys.path.insert(0, '.')
import interfaceadvection.jl_b200 as ia
n = 512
N = (n,) * 3
cen = torch.tensor([n / 2, n / 2, n / 4], device="cuda")
sim = ia.TwoPhaseSimulation(N, (0, 0, 0), float(n), T=torch.float32, lam_rho=1e-3, InterfaceSDF=lambda x: n / 8 - ((x - cen.to(x.dtype)) ** 2).sum(-1).sqrt(), perdir=(1, 2))
sim.flow.u.uniform_(-0.3, 0.3)
def t(fn, k=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
out = {"grid": list(N), "dtype": "f32"}
out["metrics_ms"] = t(lambda: ia.metrics(sim.flow.u, sim.intf.f, 1e-3, None, (0, 0, -1.0), (0, 0, 128.0)))
out["metrics_GBs"] = n ** 3 * 4 * 4 / out["metrics_ms"] / 1e6
ls = ia.LevelSet(sim)
out["computeL_ms"] = t(lambda: ia.computeL(ls.L, ls.phi, ls.phi_ini, (1, 2)))
out["computeL_GBs"] = n ** 3 * 3 * 4 / out["computeL_ms"] / 1e6
out["redistaning_d5_dtau05_ms"] = t(lambda: ia.redistaning(ls, 5, 0.5, (1, 2)), k=2)
out["redistaning_pseudo_steps"] = 10
print(json.dumps(out))
PY
cat $O/r2_s30_post_timing.json
