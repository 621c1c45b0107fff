import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import functional as F, capi
from kalman_vae_b200.functional import Problem
from kalman_vae_b200.engine import KalmanStep
from kalman_vae_b200.synthetic import Shape, make_case
dev = torch.device("cuda:0")
B, T = int(sys.argv[1]), int(sys.argv[2])
shape = Shape(B, T, 4, 2, 4, 3)
case = make_case(shape, seed=1)
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
for lanes in ([int(sys.argv[3])] if len(sys.argv) > 3 else (4, 2, 1)):
    pb = Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=lanes)
    ks = KalmanStep(pb, g["eps"], use_graphs=False)
    from kalman_vae_b200 import capi as _c
    tf = t(lambda: _c.filter_smooth_fwd(ks.pb.dims, ks._inputs, ks._states, ks.A_list, ks.B_list, ks.C_list, ks.info, ks.dev))
    ws_elbo = torch.empty(max(_c.elbo_workspace_bytes(ks.pb.dims), 16), dtype=torch.uint8, device=ks.dev)
    te = t(lambda: _c.elbo_fwd(ks.pb.dims, ks._inputs, ks._states, ks.eps, ks.jitter, ks.terms, ws_elbo, ks.info, ks.dev))
    tb = t(lambda: _c.bwd(ks.dims_bwd, ks._inputs, ks._states, ks.eps, ks.jitter, ks.g_elbo, ks.terms, None, ks.grads, ks.ws_bwd, ks.info, ks.dev))
    n = B * T
    print(f"B={B} T={T} L={lanes}: fwd {tf*1e3:.2f} ms ({440*n/tf/1e9:.0f} GB/s)  elbo {te*1e3:.2f} ms  bwd {tb*1e3:.2f} ms ({316*n/tb/1e9:.0f} GB/s)  "
          f"fused step (fwd + bwd incl. ELBO value) {n/(tf+tb)/1e9:.3f} G seq-steps/s = {772*n/(tf+tb)/1e9:.0f} GB/s = {772*n/(tf+tb)/1e9/6452.5*100:.1f}% of HBM peak", flush=True)
    del ks
    torch.cuda.empty_cache()
