import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import KalmanFilter
from kalman_vae_b200.dyn_param import PrecomputedWeights
from kalman_vae_b200.synthetic import CONFIGS, make_case
dev = torch.device("cuda:0")
shape = CONFIGS["cfg2"]
case = make_case(shape, seed=1)
dyn = PrecomputedWeights(case["A"], case["B"], case["C"])
kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, case["mu0"], case["Sigma0"], dyn).to(dev)
kf.strict = False; kf.check_info = False
host = {k: case[k].pin_memory() for k in ("Y", "U", "mask", "alpha", "eps")}
params = list(dyn.parameters())
out_host = torch.empty(1 + sum(p.numel() for p in params), pin_memory=True)
def step():
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    Y = d["Y"].requires_grad_(True)
    dyn.set_weights(d["alpha"].requires_grad_(True))
    kf._draw_eps = lambda B, T, n, like: d["eps"]
    outs = kf.smooth(Y, d["U"], d["mask"])
    val = kf.elbo(outs[0], outs[1], Y, d["U"], outs[6], outs[7], outs[8], mask=d["mask"])
    grads = torch.autograd.grad(val, [Y, dyn.alpha] + params)
    flat = torch.cat([val.detach().reshape(1)] + [g.reshape(-1) for g in grads[2:]])
    out_host.copy_(flat, non_blocking=True)
for _ in range(30): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host enqueue ms/step %.3f  incl sync %.3f" % ((t1 - t0) / 200 * 1e3, (t2 - t0) / 200 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
