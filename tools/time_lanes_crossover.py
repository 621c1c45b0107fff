"""Where does the thread-per-sequence family (lanes = 1) overtake the lane-group family (lanes = 4) for n = 4?
Times the forward launch and the fused step (CUDA graph) for both at several batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200.engine import KalmanStep
from kalman_vae_b200.functional import Problem
from kalman_vae_b200.synthetic import Shape, make_case

dev = torch.device("cuda:0")


def t(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for T in (20, 200):
    for B in (8192, 12288, 16384, 24576, 32768, 49152):
        if T == 200 and B > 32768:
            continue
        case = make_case(Shape(B, T, 4, 2, 4, 3), seed=3, mask_kind="bernoulli" if T == 200 else "ones")
        g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
        row = []
        for lanes in (4, 1):
            pb = Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=lanes)
            ks = KalmanStep(pb, g["eps"], use_graphs=True)
            reps = 50 if T == 20 else 10
            row.append((t(ks.forward_only, reps), t(ks.step, reps)))
            del ks
        print(f"T={T} B={B}: forward L4 {row[0][0]:8.1f} us  L1 {row[1][0]:8.1f} us | fwd+elbo+bwd L4 {row[0][1]:8.1f} us  L1 {row[1][1]:8.1f} us", flush=True)
        del g
        torch.cuda.empty_cache()
