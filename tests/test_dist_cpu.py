"""CPU test of the N>1 path (gloo, world_size 2): sharding + the one all-reduce reproduce the unsharded
ELBO (global mask normalisation) and parameter gradients.  The per-shard arithmetic comes from the oracle
here (no GPU); the CUDA path plugs the same two helpers in (bench.py / KalmanStep)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kalman_vae_b200.dist import allreduce_param_grads, globalize_elbo_terms, shard_bounds, shard_case
from kalman_vae_b200.synthetic import Shape, make_case


def test_shard_bounds_cover_batch():
    for B in (1, 7, 8, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kalman_oracle as ko
        torch.set_num_threads(1)
        case = make_case(Shape(7, 9, 4, 2, 4, 3), seed=5, mask_kind="bernoulli", zero_u=False, c_std=0.3)
        local = shard_case(case, rank, world)
        f = lambda k: local[k].double().clone()
        A, Bm, C = f("A").requires_grad_(), f("B").requires_grad_(), f("C").requires_grad_()
        outs, Q_seq = ko.smooth(f("Y"), f("U"), f("mask"), f("alpha"), A, Bm, C, f("Q"), f("R"), f("mu0"), f("Sigma0"), False, False)
        _, t = ko.elbo(outs[0], outs[1], f("Y"), f("U"), outs[6], outs[7], outs[8], Q_seq, f("R"), f("mu0"), f("Sigma0"),
                       f("mask"), f("eps"), return_terms=True)
        local_sum = t["trans"] + t["emiss"] + t["init"] + t["entropy"]
        terms = torch.zeros(8, dtype=torch.float64)
        terms[0], terms[1], terms[2], terms[3] = t["trans"].detach(), t["emiss"].detach(), t["init"].detach(), t["entropy"].detach()
        terms[4] = f("mask").sum()
        globalize_elbo_terms(terms)
        # local adjoint with the GLOBAL normaliser as upstream factor, then one all-reduce
        grads = list(torch.autograd.grad(local_sum * terms[6], [A, Bm, C]))
        allreduce_param_grads(grads)
        if rank == 0:
            ret["elbo"] = float(terms[5])
            ret["grads"] = [g.clone() for g in grads]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharding_matches_unsharded():
    from oracle import kalman_oracle as ko
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    case = make_case(Shape(7, 9, 4, 2, 4, 3), seed=5, mask_kind="bernoulli", zero_u=False, c_std=0.3)
    full = ko.run_case(case, torch.float64)
    assert abs(ret["elbo"] - float(full["elbo"])) < 1e-10 * abs(float(full["elbo"]))
    for g, k in zip(ret["grads"], ("dA", "dB", "dC")):
        assert float((g - full[k]).norm() / full[k].norm()) < 1e-10, k
