"""Host-side profile (cProfile) of the reference-shaped route KalmanFilter.smooth -> .elbo -> autograd.grad at cfg2."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import KalmanFilter
from kalman_vae_b200.dyn_param import PrecomputedWeights
from kalman_vae_b200.synthetic import CONFIGS, make_case
dev = torch.device("cuda:0")
shape = CONFIGS["cfg2"]
case = make_case(shape, seed=1)
d = {k: case[k].to(dev) for k in ("Y", "U", "mask", "alpha", "eps")}
dyn = PrecomputedWeights(case["A"], case["B"], case["C"], None, switching=False)
kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, case["mu0"], case["Sigma0"], dyn).to(dev)
params = list(dyn.parameters())
def step():
    Y = d["Y"].requires_grad_(True)
    dyn.set_weights(d["alpha"].requires_grad_(True))
    kf._draw_eps = lambda B, T, n, like: d["eps"]
    outs = kf.smooth(Y, d["U"], d["mask"])
    val = kf.elbo(outs[0], outs[1], Y, d["U"], outs[6], outs[7], outs[8], mask=d["mask"])
    g = torch.autograd.grad(val, [Y, dyn.alpha] + params)
    return val
for _ in range(50): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(500): step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue time per step {t_host/500*1e3:.3f} ms; incl. final sync {t_all/500*1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
