"""How much of the cfg2 adjoint is the (time-parallel) ELBO part?  Times, per lane count:
   full fused adjoint (ELBO value + ELBO adjoint + smoother/filter adjoint),
   ELBO-only adjoint (serial in t today),
   smoother+filter adjoint alone fed by dense cotangents of (mu_s, Sigma_s)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import functional as F, capi
from kalman_vae_b200.functional import Problem
from kalman_vae_b200.synthetic import CONFIGS, make_case

dev = torch.device("cuda:0")
shape = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
case = make_case(shape, seed=1)
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}


def t(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


one = torch.ones(1, device=dev)
for lanes in (4, 2, 1):
    pb = Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=lanes)
    B, T, n, p, m, K = pb.shape
    st, Al, Bl, Cl = F.smooth_fwd(pb)
    pb.mask_partials = st.mask_partials
    eps = torch.randn(B, T, n, device=dev)
    terms = torch.empty(8, device=dev)
    full = t(lambda: F.adjoint(pb, st, eps=eps, g_elbo=one, terms=terms, with_elbo=True))
    F.elbo_terms(pb, st, eps)
    eo = t(lambda: F.adjoint(pb, st, eps=eps, g_elbo=one, terms=terms, elbo_only=True))
    gr = F.adjoint(pb, st, eps=eps, g_elbo=one, terms=terms, elbo_only=True)
    cot = dict(mus_smooth=gr["dmus"].reshape(B, T, n, 1), Sigmas_smooth=gr["dSigmas"])
    sf = t(lambda: F.adjoint(pb, st, cot=cot))
    ev = t(lambda: F.elbo_terms(pb, st, eps))
    print(f"L={lanes}: full fused {full:.1f} us | elbo-only adjoint {eo:.1f} us | smoother+filter adjoint (dense cot) {sf:.1f} us | elbo value {ev:.1f} us", flush=True)
