"""Dev check (GPU box): forward kernel vs oracle for every instantiated lane count + quick timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import capi
from kalman_vae_b200.synthetic import Shape, make_case
from oracle import kalman_oracle as ko

dev = torch.device("cuda:0")
def rel(a, b): return ((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

def run_fwd(case, shape, lanes, smooth=True, lists=True):
    g = {k: (v.to(dev).contiguous() if torch.is_tensor(v) else v) for k, v in case.items()}
    B, T, n, p, m, K = shape.B, shape.T, shape.n, shape.p, shape.m, shape.K
    z = lambda *s: torch.full(s, float("nan"), device=dev)
    o = dict(mus_filt=z(B, T, n, 1), Sigmas_filt=z(B, T, n, n), mus_pred=z(B, T, n, 1), Sigmas_pred=z(B, T, n, n),
             mus_smooth=z(B, T, n, 1), Sigmas_smooth=z(B, T, n, n))
    if lists:
        o.update(A_list=z(B, T, n, n), B_list=z(B, T, n, m), C_list=z(B, T, p, n))
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    dims = capi.make_dims(B, T, n, p, m, K, shape.q_per_mode, shape.c_shared, lanes)
    assert capi.supported(dims), (shape, lanes)
    inp = capi.make_inputs(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"])
    st = capi.make_states(o["mus_filt"], o["Sigmas_filt"], o["mus_pred"], o["Sigmas_pred"],
                          o["mus_smooth"] if smooth else None, o["Sigmas_smooth"] if smooth else None)
    fn = lambda: capi.filter_smooth_fwd(dims, inp, st, o.get("A_list"), o.get("B_list"), o.get("C_list"), info, dev)
    fn(); torch.cuda.synchronize()
    o["info"] = info
    return o, fn

cases = [
    ("lstm", Shape(37, 20, 4, 2, 4, 3), dict(mask_kind="bernoulli", zero_u=False, c_std=0.3), (1, 2, 4)),
    ("switch", Shape(37, 20, 4, 2, 4, 3, True, True), dict(mask_kind="block", zero_u=False, c_std=0.3, nonsym_q=True), (1, 2, 4)),
    ("k1", Shape(5, 9, 4, 2, 4, 1), dict(mask_kind="fractional", zero_u=False, c_std=0.3), (1, 2, 4)),
    ("rocket", Shape(3, 50, 2, 1, 1, 1, True, True), dict(mask_kind="ones", zero_u=False, c_std=0.3), (1, 2)),
    ("n8", Shape(9, 15, 8, 4, 8, 4, True, True), dict(mask_kind="bernoulli", zero_u=False, c_std=0.3, nonsym_q=True), (4, 8)),
    ("n16", Shape(7, 30, 16, 8, 16, 8, True, True), dict(mask_kind="bernoulli", zero_u=False, c_std=0.3, nonsym_q=True), (8, 16)),
]
worst = 0.0
for name, shape, kw, lanes_list in cases:
    case = make_case(shape, seed=3, **kw)
    r32 = ko.run_case(case, torch.float32, want_grads=False)
    r64 = ko.run_case(case, torch.float64, want_grads=False)
    for lanes in lanes_list:
        o, _ = run_fwd(case, shape, lanes)
        errs = {}
        for k in ko.OUT_NAMES:
            e32, e64, floor = rel(o[k], r32[k]), rel(o[k], r64[k]), rel(r32[k], r64[k])
            errs[k] = f"{e32:.1e}/{e64:.1e}|{floor:.1e}"
            worst = max(worst, e64 / max(floor, 1e-7))
        print(name, "L=%d" % lanes, "info=%d" % int(o["info"]), errs, flush=True)
print("worst e64/floor ratio", worst)

# timing
for name, shape, lanes_list in [("cfg2", Shape(8192, 20, 4, 2, 4, 3), (1, 2, 4)),
                                ("cfg3/64", Shape(65536, 64, 4, 2, 4, 3), (1, 2, 4)),
                                ("cfg4/8", Shape(2048, 200, 16, 8, 16, 8, True, True), (8, 16))]:
    case = make_case(shape, seed=1)
    for lanes in lanes_list:
        o, fn = run_fwd(case, shape, lanes)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        n, p, m, K = shape.n, shape.p, shape.m, shape.K
        c = 0 if shape.c_shared else 1
        by = 4 * (p + m + 1 + K + 3 * n + 4 * n * n + n * m + c * p * n) * shape.B * shape.T
        print(f"{name} L={lanes}: {ms*1e3:.1f} us  {shape.B*shape.T/ms/1e6:.2f} G seq-steps/s  {by/ms/1e6:.0f} GB/s algorithmic", flush=True)
