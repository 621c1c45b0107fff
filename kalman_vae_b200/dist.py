"""Data parallelism for the Kalman hot path: sequences are independent, so the batch is cut into
contiguous per-rank shards (one process per GPU) and the ONLY exchange is one all-reduce of
  (a) the five ELBO partial sums [trans, emiss, init, entropy, sum(mask)] — the reference normalises
      by the GLOBAL number of observed frames (kalman_filter.py:392-400), and
  (b) the flat parameter-gradient buffer (dA,dB,dC[,dQ]; a few hundred KB, latency bound).
dY / dU / dalpha stay on the rank that owns the sequences.  Forward-only use (imputation) needs no
collective at all.

Two implementations of that one exchange:
  * PeerExchange (default on GPUs, used by engine.KalmanStep): the library's own kernels over NVLink peer memory
    (csrc/kvae_dp.cu, kvae_kf_bwd_dp) -- torch.distributed only all-gathers the CUDA-IPC handles once at set-up;
  * globalize_elbo_terms / allreduce_param_grads: torch.distributed all-reduces (NCCL on the GPUs, gloo in the CPU
    tests), for the autograd route and as the KVAE_DP_COLLECTIVE=nccl comparison path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(B: int, rank: int, world: int):
    """Contiguous [lo, hi) of the batch owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_case(case: dict, rank: int, world: int):
    """Slices every per-sequence tensor of a case dict ([B,...]) to this rank's shard."""
    B = case["Y"].shape[0]
    lo, hi = shard_bounds(B, rank, world)
    out = {}
    for k, v in case.items():
        if torch.is_tensor(v) and v.dim() >= 2 and v.shape[0] == B and k in ("Y", "U", "mask", "alpha", "eps"):
            out[k] = v[lo:hi].contiguous()
        else:
            out[k] = v
    return out


def _active(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def globalize_elbo_terms(terms: torch.Tensor, group=None) -> torch.Tensor:
    """terms[0:5] are local sums; after the all-reduce terms[5] = global elbo and terms[6] = 1/max(sum mask,1)
    (the factor the adjoint kernel multiplies the upstream gradient with).  In place; returns terms."""
    if _active(group):
        dist.all_reduce(terms[:5], op=dist.ReduceOp.SUM, group=group)
    n = terms[4].clamp(min=1.0)
    terms[5] = (terms[0] + terms[1] + terms[2] + terms[3]) / n
    terms[6] = 1.0 / n
    return terms


def allreduce_param_grads(grads, group=None):
    """Sums the parameter gradients over ranks through ONE flat buffer.  grads: list of tensors (in place)."""
    grads = [g for g in grads if g is not None]
    if not grads or not _active(group):
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[o:o + n].view_as(g))
        o += n
    return grads


class PeerExchange:
    """The data-parallel exchange of the step over NVLink peer memory (csrc/kvae_dp.cu): every rank exports a small
    cudaMalloc'ed exchange buffer through CUDA IPC, the handles are all-gathered ONCE with torch.distributed, and from
    then on kvae_dp_finalize sums [parameter gradients | ELBO sums] over the ranks, applies the global normaliser and
    scales the local per-step gradients in its own kernels: no NCCL call on the hot path."""

    def __init__(self, device, nfloats, group=None):
        from . import capi
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        # every rank takes part in BOTH gathers whatever happens locally, so a failure on one rank cannot hang the others
        self.comm, handle, ok, err = None, b"", 1, ""
        try:
            self.comm, handle = capi.dp_create(device, self.rank, self.world, nfloats)
        except capi.KvaeError as e:
            ok, err = 0, str(e)
        infos = [None] * self.world
        dist.all_gather_object(infos, (ok, err, handle), group=group)
        if all(o for o, _, _ in infos):
            try:
                capi.dp_connect(self.comm, [h for _, _, h in infos])
            except capi.KvaeError as e:   # e.g. no peer access between the devices
                ok, err = 0, str(e)
        else:
            ok = 0
            err = err or "; ".join(e for o, e, _ in infos if not o)
        oks = [None] * self.world
        dist.all_gather_object(oks, (ok, err), group=group)
        if not all(o for o, _ in oks):
            if self.comm is not None:
                capi.dp_destroy(self.comm)
                self.comm = None
            raise RuntimeError("peer-memory exchange unavailable: " + "; ".join(sorted({e for o, e in oks if not o and e})))

    def close(self):
        from . import capi
        if self.comm is not None:
            capi.dp_destroy(self.comm)
            self.comm = None
