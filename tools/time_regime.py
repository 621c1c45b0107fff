"""Times the regime-sampler kernels (fwd + adjoint) at cfg2 / cfg4 shapes against the reference's per-step torch loop
run on the same GPU (the loop the drop-in replaces: switch_dyn_param.py:67-79), CUDA events, median of 50."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200.functional import RegimeSampleFunction

dev = torch.device("cuda:0")


def torch_loop(logits, init, gumbel, trans, tau, hard):
    """the reference's op sequence (switch_dyn_param.py:51-79) with the noise factored out, on the GPU"""
    B, T, K, _ = logits.shape
    gs = lambda l, g: ((l + g) / tau).softmax(-1)
    y0 = gs(init, gumbel[:, 0])
    ys = [y0]; lq = [(y0 * torch.log_softmax(init, -1)).sum(-1)]; lp = [(y0 * torch.full_like(init, 1.0 / K).log()).sum(-1)]
    yp = y0
    for t in range(1, T):
        l_t = torch.matmul(yp.unsqueeze(1), logits[:, t]).squeeze(1)
        y_t = gs(l_t, gumbel[:, t])
        lq.append((y_t * torch.log_softmax(l_t, -1)).sum(-1))
        tp = torch.matmul(yp.unsqueeze(1), trans).squeeze(1)
        lp.append((y_t * torch.log(tp.clamp_min(1e-8))).sum(-1))
        ys.append(y_t); yp = y_t
    return torch.stack(ys, 1), torch.stack(lq, 1), torch.stack(lp, 1)


def timed(fn, reps=10, inner=20):
    """median over `reps` of (time of `inner` back-to-back calls) / inner: the queue stays full, so host-side call
    overhead is hidden behind the previous launch"""
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    return statistics.median(ts) * 1e3


for name, (B, T, K) in {"cfg2 (B=8192,T=20,K=3)": (8192, 20, 3), "cfg4 (B=16384,T=200,K=8)": (16384, 200, 8)}.items():
    g = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn(B, T, K, K, device=dev, generator=g).requires_grad_(True)
    init = torch.randn(B, K, device=dev, generator=g).requires_grad_(True)
    gumbel = -torch.empty(B, T, K, device=dev).exponential_(generator=g).log()
    trans = torch.full((K, K), 0.1 / (K - 1), device=dev); trans.fill_diagonal_(0.9)
    cy, cq, cp = torch.randn(B, T, K, device=dev), torch.randn(B, T, device=dev), torch.randn(B, T, device=dev)

    def run(f):
        y, lq, lp = f(logits, init, gumbel, trans, 0.5, False)
        loss = (cy * y).sum() + (cq * lq).sum() + (cp * lp).sum()
        return torch.autograd.grad(loss, [logits, init])

    from kalman_vae_b200 import capi
    y_b, lq_b, lp_b = (torch.empty(B, T, K, device=dev), torch.empty(B, T, device=dev), torch.empty(B, T, device=dev))
    dl_b, di_b = torch.empty(B, T, K, K, device=dev), torch.empty(B, K, device=dev)
    lg_d, in_d = logits.detach(), init.detach()
    kf_us = timed(lambda: capi.regime_fwd(B, T, K, False, 0.5, lg_d, in_d, gumbel, trans, y_b, lq_b, lp_b, dev))
    kb_us = timed(lambda: capi.regime_bwd(B, T, K, False, 0.5, lg_d, in_d, gumbel, trans, y_b, cy, cq, cp, dl_b, di_b, dev))
    print(f"{name}: k_regime_fwd {kf_us:.1f} us, k_regime_bwd {kb_us:.1f} us (C-ABI calls, CUDA events incl. launch)", flush=True)
    k_us = timed(lambda: run(RegimeSampleFunction.apply))
    t_us = timed(lambda: run(torch_loop), reps=3, inner=2)
    ga, gb = run(RegimeSampleFunction.apply), run(torch_loop)
    err = max(float((x - y).norm() / y.norm()) for x, y in zip(ga, gb))
    bytes_alg = 4 * B * T * (K * K * 2 + K * 3 + 4)      # logits r + d_logits w, gumbel r, y w+r, cot, log_q/p + cots
    print(f"{name}: autograd route fwd+bwd incl. the 6 loss ops {k_us:.1f} us, torch per-step loop on the same GPU "
          f"{t_us:.0f} us ({t_us / k_us:.0f}x), grad rel diff {err:.1e}", flush=True)
