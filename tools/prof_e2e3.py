import sys, os, time, gc
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
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
def step():
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    Y = d["Y"].requires_grad_(True)
    dyn.set_weights(d["alpha"].requires_grad_(True))
    kf._draw_eps = lambda B, T, n, like: d["eps"]
    t = time.perf_counter()
    outs = kf.smooth(Y, d["U"], d["mask"])
    ts = time.perf_counter() - t
    if mode == "smooth": return ts
    val = kf.elbo(outs[0], outs[1], Y, d["U"], outs[6], outs[7], outs[8], mask=d["mask"])
    grads = torch.autograd.grad(val, [Y, dyn.alpha] + params)
    return ts
for i in range(40):
    ts = step()
    if i % 4 == 0:
        st = torch.cuda.memory_stats()
        print(i, "smooth ms %.2f" % (ts * 1e3), "alloc MB", torch.cuda.memory_allocated() >> 20, "reserved MB", torch.cuda.memory_reserved() >> 20,
              "dev_allocs", st["num_device_alloc"], "gc", gc.get_count(), flush=True)
import itertools
def timed(label, fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(label, "enqueue ms/step %.3f  total ms/step %.3f" % ((t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3), flush=True)
mode = "full"
timed("full step (pinned h2d)", step)
def h2d_only():
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
timed("h2d only", h2d_only)
dres = {k: v.to(dev) for k, v in host.items()}
def step_resident():
    Y = dres["Y"].detach().requires_grad_(True)
    dyn.set_weights(dres["alpha"].detach().requires_grad_(True))
    kf._draw_eps = lambda B, T, n, like: dres["eps"]
    outs = kf.smooth(Y, dres["U"], dres["mask"])
    val = kf.elbo(outs[0], outs[1], Y, dres["U"], outs[6], outs[7], outs[8], mask=dres["mask"])
    grads = torch.autograd.grad(val, [Y, dyn.alpha] + params)
timed("full step (device-resident inputs)", step_resident)
def step_d2h():
    step_resident()
    out_host[:1].copy_(torch.ones(1, device=dev), non_blocking=True)
timed("resident + d2h", step_d2h)
