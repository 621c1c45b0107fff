"""GPU-box validation at BASELINE.json sizes: cfg3 (imputation, T=1000, masked, forward only) and cfg4
(SKVAE n=16, fwd+bwd).  No oracle run at these sizes: size-independent properties + throughput."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import functional as F, capi
from kalman_vae_b200.functional import Problem
from kalman_vae_b200.synthetic import Shape, make_case, CONFIGS

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
Bover = int(sys.argv[2]) if len(sys.argv) > 2 else 0

def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3

if which == "cfg3":
    shape = CONFIGS["cfg3"]
    if Bover: shape = Shape(Bover, shape.T, shape.n, shape.p, shape.m, shape.K)
    for mask_kind in ("block", "bernoulli"):
        case = make_case(shape, seed=10, mask_kind=mask_kind)
        g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
        for lanes in (1, 4):
            pb = Problem(g["Y"], None, g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=lanes)
            B, T, n, p, m, K = pb.shape
            e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
            st = F.States(e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n))
            Al, Bl, Cl = e(B, T, n, n), e(B, T, n, m), e(B, T, p, n)
            info = F.info_word(dev); info.zero_()
            fn = lambda: capi.filter_smooth_fwd(pb.dims, pb.inputs(), st.c_struct(), Al, Bl, Cl, info, dev)
            sec = ev_time(fn)
            miss = g["mask"] == 0
            ok_mu = bool(torch.equal(st.mus_filt[miss], st.mus_pred[miss]))
            sp = st.Sigmas_pred[miss]
            ok_sig = bool(torch.equal(st.Sigmas_filt[miss], 0.5 * (sp + sp.mT)))
            del sp
            finite = bool(torch.isfinite(st.Sigmas_smooth).all() and torch.isfinite(st.mus_smooth).all())
            sym = bool(torch.equal(st.Sigmas_smooth[:, :-1], st.Sigmas_smooth[:, :-1].mT))
            by = 4 * (p + m + 1 + K + 3 * n + 4 * n * n + n * m + p * n)
            r = dict(B=B, T=T, mask=mask_kind, lanes=lanes, seconds=sec, seq_steps_per_s=B * T / sec, alg_GBps=by * B * T / sec / 1e9,
                     frac_of_6452=by * B * T / sec / 1e9 / 6452.5, info=int(info), mask0_mu_bitexact=ok_mu, mask0_sigma_bitexact=ok_sig,
                     finite=finite, smooth_symmetric=sym)
            print(json.dumps(r), flush=True)
            del st, Al, Bl, Cl
            torch.cuda.empty_cache()
else:
    shape = CONFIGS["cfg4"]
    if Bover: shape = Shape(Bover, shape.T, shape.n, shape.p, shape.m, shape.K, True, True)
    case = make_case(shape, seed=10)
    g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
    for lanes in (16,):
        pb = Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], True, True, lanes=lanes)
        F.info_word(dev).zero_()
        st, *_ = F.smooth_fwd(pb)
        tf = ev_time(lambda: F.smooth_fwd(pb), 2)
        terms = F.elbo_terms(pb, st, g["eps"])
        te = ev_time(lambda: F.elbo_terms(pb, st, g["eps"]), 2)
        gel = torch.ones(1, device=dev)
        gr = F.adjoint(pb, st, eps=g["eps"], g_elbo=gel, terms=terms, need_dU=False)
        tb = ev_time(lambda: F.adjoint(pb, st, eps=g["eps"], g_elbo=gel, terms=terms, need_dU=False), 2)
        # the fused training step (k_filter_smooth, k_bwd with the ELBO value, k_bwd_final) in one CUDA graph
        from kalman_vae_b200.engine import KalmanStep
        ks = KalmanStep(pb, g["eps"], use_graphs=True, need_dU=False)
        tstep = ev_time(ks.step, 3)
        B, T = shape.B, shape.T
        finite = all(bool(torch.isfinite(v).all()) for v in gr.values() if v is not None)
        print(json.dumps(dict(B=B, T=T, lanes=lanes, fwd_s=tf, elbo_s=te, bwd_s=tb, seq_steps_per_s=B * T / (tf + te + tb),
                              alg_GBps=9032 * B * T / (tf + te + tb) / 1e9, fused_step_s=tstep, fused_seq_steps_per_s=B * T / tstep,
                              fused_elbo=float(ks.terms[5]), elbo=float(terms[5]), info=int(F.info_word(dev)), finite=finite)), flush=True)
