import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import functional as F, capi
from kalman_vae_b200.functional import Problem, States
from kalman_vae_b200.synthetic import CONFIGS, make_case
dev = torch.device("cuda:0")
shape = CONFIGS["cfg2"]
case = make_case(shape, seed=1)
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
def t(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for lanes in (4, 2, 1):
    pb = Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=lanes)
    B, T, n, p, m, K = pb.shape
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    st = States(e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n))
    stf = States(st.mus_filt, st.Sigmas_filt, st.mus_pred, st.Sigmas_pred)
    Al, Bl, Cl = e(B, T, n, n), e(B, T, n, m), e(B, T, p, n)
    info = F.info_word(dev)
    inp = pb.inputs()
    f_only = t(lambda: capi.filter_smooth_fwd(pb.dims, inp, stf.c_struct(), Al, Bl, Cl, info, dev))
    f_nolists = t(lambda: capi.filter_smooth_fwd(pb.dims, inp, stf.c_struct(), None, None, None, info, dev))
    fs = t(lambda: capi.filter_smooth_fwd(pb.dims, inp, st.c_struct(), Al, Bl, Cl, info, dev))
    print(f"L={lanes}: filter only {f_only:.1f} us, filter no lists {f_nolists:.1f} us, filter+smooth {fs:.1f} us", flush=True)
