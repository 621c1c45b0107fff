"""KVAE.impute's Kalman part: smooth() + the two C_t mu products as the reference does them, against
KalmanFilter.impute_observations (no A_list/B_list/C_list materialised).  CUDA events, median of 10."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import KalmanFilter
from kalman_vae_b200.dyn_param import PrecomputedWeights
from kalman_vae_b200.synthetic import Shape, make_case

dev = torch.device("cuda:0")
for B, T in ((8192, 20), (65536, 200)):
    case = make_case(Shape(B, T, 4, 2, 4, 3), seed=3, mask_kind="bernoulli")
    dyn = PrecomputedWeights(case["A"], case["B"], case["C"])
    kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, case["mu0"], case["Sigma0"], dyn).to(dev).eval()
    Y, U, mask = case["Y"].to(dev), case["U"].to(dev), case["mask"].to(dev)
    dyn.set_weights(case["alpha"].to(dev))

    def ref_style():
        o = kf.smooth(Y, U, mask)
        return (o[8] @ o[0]).squeeze(-1), (o[8] @ o[2]).squeeze(-1)

    def fused():
        return kf.impute_observations(Y, U, mask)[:2]

    def t(fn):
        with torch.no_grad():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    with torch.no_grad():
        r, f = ref_style(), fused()
    err = max(float((x - y).abs().max()) for x, y in zip(r, f))
    tr, tf = t(ref_style), t(fused)
    print(f"B={B} T={T}: smooth() + 2 bmm {tr:.3f} ms, impute_observations {tf:.3f} ms ({tr / tf:.2f}x), max abs diff {err:.1e}", flush=True)
