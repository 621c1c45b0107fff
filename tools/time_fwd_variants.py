"""Development: forward kernel with / without the list outputs and the smoother (which streams bound it?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import capi, functional as F
from kalman_vae_b200.functional import Problem, States
from kalman_vae_b200.synthetic import Shape, make_case
dev = torch.device("cuda:0")
B, T, lanes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
case = make_case(Shape(B, T, 4, 2, 4, 3), seed=1)
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
pb = Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=lanes)
n, p, m = 4, 2, 4
e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
st = States(e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n))
stf = States(st.mus_filt, st.Sigmas_filt, st.mus_pred, st.Sigmas_pred)
Al, Bl, Cl = e(B, T, n, n), e(B, T, n, m), e(B, T, p, n)
info = F.info_word(dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
inp = pb.inputs()
N = B * T
for name, fn, by in (("filter+smooth, lists", lambda: capi.filter_smooth_fwd(pb.dims, inp, st.c_struct(), Al, Bl, Cl, info, dev), 440),
                     ("filter+smooth, no lists", lambda: capi.filter_smooth_fwd(pb.dims, inp, st.c_struct(), None, None, None, info, dev), 280),
                     ("filter only, lists", lambda: capi.filter_smooth_fwd(pb.dims, inp, stf.c_struct(), Al, Bl, Cl, info, dev), 360),
                     ("filter only, no lists", lambda: capi.filter_smooth_fwd(pb.dims, inp, stf.c_struct(), None, None, None, info, dev), 200)):
    ms = t(fn)
    print(f"B={B} T={T} L={lanes} {name:26s} {ms:7.3f} ms  {by * N / ms / 1e6:7.0f} GB/s algorithmic", flush=True)
