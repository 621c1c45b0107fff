"""The reference's KVAE training step with this package's Kalman block swapped in (BASELINE.json configs[4], "cfg5").

The reference model (`kvae/model/model.py: KVAE` — conv encoder / decoder, reparameterisation, VAE loss) is used AS IS,
imported from the reference sources (the pip-installed copy under baseline/_ref, or /root/reference in the build
container); the only change is the one INTEGRATION.md section 2 describes: the three names `KalmanFilter`,
`base_dyn_param`, `switch_dyn_param` in `kvae.model.model` point at `kalman_vae_b200`.  The step body mirrors
kvae/train/train.py:32-58 (that module itself needs pytorch_lightning / imageio to import): reset_state, ones mask,
zero_grad, forward, compute_loss, backward, clip_grad_norm_(10), Adam(lr 0.007).

Used by bench.py --workload cfg5 and tests/test_gpu_kvae.py.  Not on the hot path: harness code.
"""
from __future__ import annotations

import os
import sys
import types

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOTS = (os.path.join(_REPO, "baseline", "_ref"), "/root/reference")


def reference_root():
    for r in REFERENCE_ROOTS:
        if os.path.isfile(os.path.join(r, "kvae", "model", "model.py")):
            return r
    return None


def load_reference_model_module():
    """kvae.model.model of the unmodified reference (two import shims: matplotlib is imported but unused by
    kalman_filter.py:5; losses.py:3 imports kvae.vae.config, which is kvae.utils.config)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found: run oracle/install_reference.sh (baseline/_ref)")
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if root not in sys.path:
        sys.path.insert(0, root)
    import kvae.utils.config as _cfg
    sys.modules.setdefault("kvae.vae.config", _cfg)
    import kvae.model.model as model_mod
    return model_mod, _cfg


class swapped_kalman:
    """Context manager: inside it `kvae.model.model` constructs its Kalman block from kalman_vae_b200 (the swap of
    INTEGRATION.md section 2, done by assignment instead of by editing the three import lines)."""

    def __init__(self, model_mod):
        self.m = model_mod

    def __enter__(self):
        import kalman_vae_b200
        from kalman_vae_b200 import dyn_param
        self.saved = (self.m.KalmanFilter, self.m.base_dyn_param, self.m.switch_dyn_param)
        self.m.KalmanFilter = kalman_vae_b200.KalmanFilter
        self.m.base_dyn_param = dyn_param
        self.m.switch_dyn_param = dyn_param
        return self

    def __exit__(self, *exc):
        self.m.KalmanFilter, self.m.base_dyn_param, self.m.switch_dyn_param = self.saved
        return False


def build_kvae(dynamics_model="lstm", drop_in=True, device=None, seed=10, **cfg_over):
    """The reference KVAE (default KVAEConfig: 32x32 frames, a_dim 2, z_dim 4, K 3) on `device`."""
    model_mod, cfg_mod = load_reference_model_module()
    cfg = cfg_mod.KVAEConfig(dynamics_model=dynamics_model, **cfg_over)
    torch.manual_seed(seed)
    if drop_in:
        with swapped_kalman(model_mod):
            model = model_mod.KVAE(cfg)
    else:
        model = model_mod.KVAE(cfg)
    if device is not None:
        model = model.to(device)
    return model, cfg


class ReferenceTrainStep:
    """One optimisation step of kvae/train/train.py:32-58 on synthetic bouncing-ball frames."""

    def __init__(self, device, drop_in=True, dynamics_model="lstm", batch=32, T=20, seed=10, distributed=False, lr=0.007,
                 grad_clip_norm=10.0):
        self.device, self.batch, self.T, self.distributed = device, batch, T, distributed
        self.model, self.cfg = build_kvae(dynamics_model, drop_in, device, seed)
        self.model.train()
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr)     # config.yaml:26
        self.clip = grad_clip_norm                                       # train.py:349
        self.last_loss = float("nan")
        self.last = None
        if distributed:   # same initial weights on every rank
            for p in self.model.parameters():
                torch.distributed.broadcast(p.data, src=0)

    def synthetic_batch(self, seed=0):
        """[B,T,1,32,32] frames of a white ball (radius ~3 px) bouncing in the frame, values in {0,1}."""
        from .synthetic import bouncing_ball
        gen = torch.Generator().manual_seed(seed)
        B, T = self.batch, self.T
        H = W = 32
        pos = bouncing_ball(B, T, 2, gen, noise_std=0.0)                 # [-1,1]^2
        cy = (pos[..., 0] * 0.8 + 1.0) * 0.5 * (H - 1)
        cx = (pos[..., 1] * 0.8 + 1.0) * 0.5 * (W - 1)
        yy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
        xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
        d2 = (yy - cy.view(B, T, 1, 1)) ** 2 + (xx - cx.view(B, T, 1, 1)) ** 2
        return (d2 <= 9.0).to(torch.float32).unsqueeze(2)

    def step(self, x):
        model = self.model
        model.kalman_filter.dyn_params.reset_state()                     # train.py:34
        x = x.to(self.device).float()
        B, T = x.shape[:2]
        mask = torch.ones(B, T, device=self.device, dtype=x.dtype)       # train.py:41
        self.opt.zero_grad(set_to_none=True)
        outputs = model(x, mask=mask)                                     # train.py:45
        losses = model.compute_loss(x, outputs, kf_weight=1.0, vae_weight=1.0, mask=mask)
        loss = losses["loss"]
        loss.backward()                                                   # train.py:53
        if self.distributed:
            grads = [p.grad for p in model.parameters() if p.grad is not None]
            flat = torch.cat([g.reshape(-1) for g in grads])
            torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
            flat /= torch.distributed.get_world_size()
            o = 0
            for g in grads:
                g.copy_(flat[o:o + g.numel()].view_as(g))
                o += g.numel()
        if self.clip and self.clip > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), self.clip)  # train.py:55-56
        self.opt.step()                                                   # train.py:58
        self.last = losses
        self.last_loss = loss.detach()
        return self.last_loss


class GraphedTrainStep(ReferenceTrainStep):
    """The same optimisation step captured ONCE into CUDA graphs and replayed: at batch 32 the eager step is bound by
    Python / launch overhead (~300 small kernels and three host reads per step), not by the GPU.

    What is captured: the reference model's `forward` as it is (encoder, this package's Kalman block, decoder), the loss
    composition of `KVAE.compute_loss` (model.py:189-229: vae_loss + kalman_filter.elbo; its two diagnostic host reads
    `count_active_units(...)` / `.item()` at :231-240 cannot be part of a graph and are left out), backward, gradient
    clipping and Adam (capturable=True: the same update with its step counter on the device).  The VAE-side reductions
    run in this package's kernels (kalman_vae_b200.vae_loss): the reference's vae_loss builds
    device constants from host scalars on every call (losses.py:17), which a capture cannot contain.
    Data parallel: graph 1 = forward + backward, then ONE eager NCCL all-reduce of the flattened gradients, graph 2 =
    clip + Adam.  Per step the host copies the batch into the static input buffer and replays.
    """

    def __init__(self, device, dynamics_model="lstm", batch=32, T=20, seed=10, distributed=False, lr=0.007,
                 grad_clip_norm=10.0):
        super().__init__(device, True, dynamics_model, batch, T, seed, distributed, lr, grad_clip_norm)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, capturable=True)
        self.model.kalman_filter.check_info = False      # no host reads inside a graph; call check() when convenient
        self.g_main = self.g_opt = None
        self.static_x = torch.zeros(batch, T, 1, 32, 32, device=device)
        self.mask = torch.ones(batch, T, device=device)                  # train.py:41
        self.static_loss = None
        self._flat = None

    def _loss(self):
        model, cfg = self.model, self.model.config
        x, mask = self.static_x, self.mask
        model.kalman_filter.dyn_params.reset_state()                     # train.py:34
        out = model(x, mask=mask)                                        # train.py:45
        from .vae_loss import vae_loss
        x_var = cfg.noise_pixel_var
        vae_elbo, _, _ = vae_loss(x, out.get("x_logits", out["x_recon"]), x_var, out["a_samples"], out["a_mu"], out["a_var"],
                                  scale_reconstruction=cfg.scale_reconstruction, mask=mask, out_distr=cfg.out_distr,
                                  beta=model.beta)                       # model.py:208-215
        A_list, B_list, C_list = out["ABC"]
        elbo_kf = model.kalman_filter.elbo(out["mus_smooth"], out["Sigmas_smooth"], out["a_samples"], out["u"],
                                           A_list, B_list, C_list, mask=mask)   # model.py:218-222
        return -(vae_elbo + elbo_kf)                                     # model.py:225-226 with both weights 1

    def _allreduce(self):
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        if self._flat is None:
            self._flat = torch.empty(sum(g.numel() for g in grads), device=self.device)
        torch._foreach_copy_(list(self._flat.split([g.numel() for g in grads])), [g.reshape(-1) for g in grads])
        torch.distributed.all_reduce(self._flat, op=torch.distributed.ReduceOp.SUM)
        self._flat /= torch.distributed.get_world_size()
        torch._foreach_copy_([g.reshape(-1) for g in grads], list(self._flat.split([g.numel() for g in grads])))

    def _update(self):
        if self.clip and self.clip > 0:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)   # train.py:55-56
        self.opt.step()                                                          # train.py:58

    def capture(self, x_example, warmup=3):
        dev = self.device
        self.static_x.copy_(x_example.to(dev).float())
        kf = self.model.kalman_filter
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):            # eager warm-up on a side stream (allocator, cuDNN plans, lazy inits)
            for _ in range(warmup):
                self.opt.zero_grad(set_to_none=True)
                self._loss().backward()
                if self.distributed:
                    self._allreduce()
                self._update()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        kf._poll_deferred()
        kf._deferred = []
        self.opt.zero_grad(set_to_none=True)
        self.g_main = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_main):
            self.static_loss = self._loss()
            self.static_loss.backward()
            if not self.distributed:
                self._update()
        if self.distributed:
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt, pool=self.g_main.pool()):
                self._update()
        return self

    def step(self, x):
        if self.g_main is None:
            self.capture(x)
        self.static_x.copy_(x, non_blocking=True)
        self.g_main.replay()
        if self.distributed:
            self._allreduce()
            self.g_opt.replay()
        self.last_loss = self.static_loss
        return self.static_loss

    def check(self):
        """One synchronising look at the Kalman kernels' status word (a replayed graph runs no host code, so the lazy
        checks of KalmanFilter.elbo are off): raises if a factorisation met a non-positive pivot since the last call."""
        from .functional import info_word
        w = info_word(self.device)
        code = int(w.item())
        if code:
            w.zero_()
            raise torch.linalg.LinAlgError(f"kvae: a Cholesky / LU pivot was not positive in a replayed step (status word {code})")
