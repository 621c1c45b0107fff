"""Runs the forward / backward launches a few times for one (B, T, lanes): the command ncu is pointed at.
   python tools/run_kernels_once.py B T lanes [fwd|bwd|both] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import capi
from kalman_vae_b200.functional import Problem
from kalman_vae_b200.engine import KalmanStep
from kalman_vae_b200.synthetic import Shape, make_case

B, T, lanes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
what = sys.argv[4] if len(sys.argv) > 4 else "both"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda:0")
case = make_case(Shape(B, T, 4, 2, 4, 3), seed=1)
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
pb = Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=lanes)
ks = KalmanStep(pb, g["eps"], use_graphs=False)
for _ in range(reps):
    if what in ("fwd", "both"):
        capi.filter_smooth_fwd(ks.pb.dims, ks._inputs, ks._states, ks.A_list, ks.B_list, ks.C_list, ks.info, ks.dev)
    if what in ("bwd", "both"):
        if what == "bwd" and _ == 0:
            capi.filter_smooth_fwd(ks.pb.dims, ks._inputs, ks._states, ks.A_list, ks.B_list, ks.C_list, ks.info, ks.dev)
        capi.bwd(ks.pb.dims, ks._inputs, ks._states, ks.eps, ks.jitter, ks.g_elbo, ks.terms, None, ks.grads, ks.ws_bwd, ks.info, ks.dev)
torch.cuda.synchronize()
print("info", int(ks.info), "done")
