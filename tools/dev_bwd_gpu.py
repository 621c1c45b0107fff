"""Dev check (GPU box): ELBO + explicit adjoint kernels vs oracle autograd, all lane counts + timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import functional as F
from kalman_vae_b200.functional import Problem, prep
from kalman_vae_b200.synthetic import Shape, make_case
from oracle import kalman_oracle as ko

dev = torch.device("cuda:0")
def rel(a, b): return ((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

def problem(case, lanes):
    g = {k: (v.to(dev).contiguous() if torch.is_tensor(v) else v) for k, v in case.items()}
    return Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"],
                   bool(case["q_per_mode"]), bool(case["c_shared"]), lanes=lanes), g

cases = [
    ("lstm", Shape(37, 20, 4, 2, 4, 3), dict(mask_kind="bernoulli", zero_u=False, c_std=0.3), (1, 2, 4)),
    ("switch", Shape(37, 20, 4, 2, 4, 3, True, True), dict(mask_kind="block", zero_u=False, c_std=0.3, nonsym_q=True), (1, 2, 4)),
    ("k1", Shape(5, 9, 4, 2, 4, 1), dict(mask_kind="fractional", zero_u=False, c_std=0.3), (1, 4)),
    ("rocket", Shape(3, 50, 2, 1, 1, 1, True, True), dict(mask_kind="ones", zero_u=False, c_std=0.3), (1, 2)),
    ("n8", Shape(9, 15, 8, 4, 8, 4, True, True), dict(mask_kind="bernoulli", zero_u=False, c_std=0.3, nonsym_q=True), (4, 8)),
    ("n16", Shape(7, 30, 16, 8, 16, 8, True, True), dict(mask_kind="bernoulli", zero_u=False, c_std=0.3, nonsym_q=True), (8, 16)),
]
if len(sys.argv) > 1 and sys.argv[1] == "parity" or len(sys.argv) == 1:
  for name, shape, kw, lanes_list in cases:
    case = make_case(shape, seed=3, **kw)
    B, T, n, p, m = shape.B, shape.T, shape.n, shape.p, shape.m
    gen = torch.Generator().manual_seed(5)
    shp = dict(mus_smooth=(B,T,n,1),Sigmas_smooth=(B,T,n,n),mus_filt=(B,T,n,1),Sigmas_filt=(B,T,n,n),mus_pred=(B,T,n,1),
               Sigmas_pred=(B,T,n,n),A_list=(B,T,n,n),B_list=(B,T,n,m),C_list=(B,T,p,n))
    for use_cot in (False, True):
        cot = {k: 0.1 * torch.randn(*s, generator=gen) for k, s in shp.items()} if use_cot else None
        if use_cot and shape.c_shared: cot["C_list"] = None
        r32 = ko.run_case(case, torch.float32, cotangents=cot); r64 = ko.run_case(case, torch.float64, cotangents=cot)
        for lanes in lanes_list:
            pb, g = problem(case, lanes)
            st, A_list, B_list, C_list = F.smooth_fwd(pb)
            F.info_word(dev).zero_()
            terms = F.elbo_terms(pb, st, g["eps"])
            gel = torch.ones(1, device=dev)
            cot_d = {k: v.to(dev).contiguous() for k, v in cot.items() if v is not None} if cot else None
            gr = F.adjoint(pb, st, eps=g["eps"], g_elbo=gel, terms=terms, cot=cot_d)
            torch.cuda.synchronize()
            out = dict(elbo=terms[5], dY=gr["dY"], dU=gr["dU"], dalpha=gr["dalpha"], dA=gr["dA"], dB=gr["dBm"], dC=gr["dC"])
            if shape.q_per_mode: out["dQ"] = gr["dQ"]
            print(name, "cot" if use_cot else "elbo", "L=%d" % lanes, "info=%d" % int(F.info_word(dev)),
                  {k: f"{rel(v, r32[k]):.1e}/{rel(v, r64[k]):.1e}|{rel(r32[k], r64[k]):.1e}" for k, v in out.items()}, flush=True)

# timing: fwd, elbo, bwd separately
for name, shape, lanes_list in [("cfg2", Shape(8192, 20, 4, 2, 4, 3), (1, 2, 4)),
                                ("cfg2sw", Shape(8192, 20, 4, 2, 4, 3, True, True), (4,)),
                                ("cfg4/16", Shape(1024, 200, 16, 8, 16, 8, True, True), (8, 16))]:
    case = make_case(shape, seed=1)
    for lanes in lanes_list:
        pb, g = problem(case, lanes)
        st, *_ = F.smooth_fwd(pb)
        terms = F.elbo_terms(pb, st, g["eps"])
        gel = torch.ones(1, device=dev)
        def t(fn, reps=10):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps * 1e3
        tf = t(lambda: F.smooth_fwd(pb))
        te = t(lambda: F.elbo_terms(pb, st, g["eps"]))
        tb = t(lambda: F.adjoint(pb, st, eps=g["eps"], g_elbo=gel, terms=terms))
        tot = tf + te + tb
        print(f"{name} L={lanes}: fwd {tf:.1f} us  elbo {te:.1f} us  bwd {tb:.1f} us  total {tot:.1f} us  "
              f"{shape.B*shape.T/tot:.2f} M seq-steps/s/us -> {shape.B*shape.T/tot*1e-3:.3f} G/s", flush=True)
