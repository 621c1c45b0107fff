import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import capi
from kalman_vae_b200.functional import Problem
from kalman_vae_b200.engine import KalmanStep
from kalman_vae_b200.synthetic import Shape, make_case
dev = torch.device("cuda:0")
B, T = int(sys.argv[1]), int(sys.argv[2])
case = make_case(Shape(B, T, 4, 2, 4, 3), seed=77)
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
def ev(fn, reps, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for graphs in (False, True):
    pb = Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=0)
    ks = KalmanStep(pb, g["eps"], use_graphs=graphs, need_dU=False)
    print("lanes", capi.pick_lanes(pb.dims), "graphs", graphs)
    for reps in (5, 20, 50):
        print(f"  fwd reps={reps}: {ev(ks.forward_only, reps):.3f} ms   step: {ev(ks.step, reps):.3f} ms", flush=True)
    del ks
