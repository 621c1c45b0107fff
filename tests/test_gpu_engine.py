"""GPU tests of KalmanStep (pre-planned step, CUDA graphs) and of the data-parallel path over NCCL."""
import os

import pytest
import torch

from kalman_vae_b200 import functional as F
from kalman_vae_b200.engine import KalmanStep
from kalman_vae_b200.functional import Problem
from kalman_vae_b200.synthetic import Shape, make_case

pytestmark = pytest.mark.gpu


def _problem(case, dev, lanes=0):
    g = {k: (v.to(dev).float().contiguous() if torch.is_tensor(v) else v) for k, v in case.items()}
    return Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"],
                   bool(case["q_per_mode"]), bool(case["c_shared"]), lanes=lanes), g


@pytest.mark.parametrize("graphs", [False, True])
def test_step_equals_functional_path(graphs):
    dev = torch.device("cuda:0")
    case = make_case(Shape(300, 12, 4, 2, 4, 3), seed=8, mask_kind="bernoulli", zero_u=False, c_std=0.3)
    pb, g = _problem(case, dev)
    st, *_ = F.smooth_fwd(pb)
    terms = F.elbo_terms(pb, st, g["eps"])
    ref = F.adjoint(pb, st, eps=g["eps"], g_elbo=torch.ones(1, device=dev), terms=terms, need_dU=False)
    t_f = torch.empty(8, device=dev)
    fused = F.adjoint(pb, st, eps=g["eps"], g_elbo=torch.ones(1, device=dev), terms=t_f, need_dU=False, with_elbo=True)
    ks = KalmanStep(pb, g["eps"], use_graphs=graphs)
    for _ in range(3):                       # replaying must be idempotent
        t = ks.step()
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    assert torch.equal(t[:7], t_f[:7])
    assert float(t[4]) == float(terms[4])
    assert rel(t[:7], terms[:7]) < 1e-6            # fused value (accumulated by the adjoint sweep) vs the ELBO kernel
    for k in ("dY", "dalpha", "dA", "dBm", "dC"):
        assert torch.equal(ks.grads[k], fused[k]), k   # same kernels, deterministic reductions -> bit-identical
        # two-call sequence: the normaliser enters inside the sweep instead of after it -> every product rounds
        # differently; dY is a cancelling sum whose fp32 noise floor is ~1e-5 (tests/_util.py)
        assert rel(ks.grads[k], ref[k]) < 5e-5, k


def test_host_pipeline_matches_device_resident_step():
    """HostPipeline: pinned host inputs, copy stream / compute stream, rotating slots -> same numbers as KalmanStep."""
    from kalman_vae_b200.engine import HostPipeline
    dev = torch.device("cuda:0")
    shape = Shape(257, 9, 4, 2, 4, 3)
    cases = [make_case(shape, seed=20 + i, mask_kind="bernoulli", zero_u=False, c_std=0.3) for i in range(5)]
    params = {k: cases[0][k].to(dev).float().contiguous() for k in ("A", "B", "C", "Q", "R", "mu0", "Sigma0")}
    pipe = HostPipeline((shape.B, shape.T, shape.n, shape.p, shape.m, shape.K), params, device=dev)
    got, slots = [], []
    hosts = [{k: c[k].float().contiguous().pin_memory() for k in ("Y", "U", "mask", "alpha", "eps")} for c in cases]
    for i, h in enumerate(hosts):
        k = pipe.step(h["Y"], h["U"], h["mask"], h["alpha"], h["eps"])
        if i >= 1:    # read the PREVIOUS step's result while this one is in flight
            elbo, flat = pipe.result(slots[-1])
            got.append((elbo, flat.clone()))
        slots.append(k)
    elbo, flat = pipe.result(slots[-1])
    got.append((elbo, flat.clone()))
    dY_last = pipe.device_grads(slots[-1])["dY"].clone()
    torch.cuda.synchronize()
    for i, c in enumerate(cases):
        c2 = dict(c)
        for k in ("A", "B", "C", "Q", "R", "mu0", "Sigma0"):
            c2[k] = cases[0][k]
        pb, g = _problem(c2, dev)
        ks = KalmanStep(pb, g["eps"], use_graphs=False)
        t = ks.step()
        torch.cuda.synchronize()
        assert got[i][0] == float(t[5]), i
        assert torch.equal(got[i][1], ks.flat.cpu()), i
        if i == len(cases) - 1:
            assert torch.equal(dY_last, ks.grads["dY"])


def _dp_worker(rank, world, port, ret, collective):
    import torch.distributed as dist
    os.environ["KVAE_DP_COLLECTIVE"] = collective
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from kalman_vae_b200.dist import shard_case
        case = make_case(Shape(301, 12, 4, 2, 4, 3), seed=9, mask_kind="bernoulli", zero_u=False, c_std=0.3)
        pb, g = _problem(shard_case(case, rank, world), dev)
        ks = KalmanStep(pb, g["eps"], use_graphs=True)
        assert ks.collective == ("nvlink-peer-memory" if collective.startswith("peer") else "nccl"), ks.collective
        for _ in range(5):                       # replays: the peer exchange alternates its two slots
            t = ks.step()
        torch.cuda.synchronize()
        assert int(ks.info) == 0
        ret[rank] = dict(elbo=float(t[5]), dA=ks.grads["dA"].cpu(), dC=ks.grads["dC"].cpu(), dY=ks.grads["dY"].cpu(),
                         dalpha=ks.grads["dalpha"].cpu())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("collective", ["peer", "peer2", "nccl"])
def test_two_gpu_data_parallel_matches_single_gpu(collective):
    """collective = "peer": the exchange over NVLink peer memory fused into the adjoint's final kernel (kvae_kf_bwd_dp);
    "peer2": the same exchange as two extra launches (kvae_kf_bwd + kvae_dp_finalize); "nccl": torch.distributed
    all-reduce + scaling kernels.  All must reproduce the single-GPU step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from kalman_vae_b200.dist import shard_bounds
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, 29600 + os.getpid() % 300 + {"peer": 7, "peer2": 13}.get(collective, 0), ret, collective), nprocs=2, join=True)
    dev = torch.device("cuda:0")
    case = make_case(Shape(301, 12, 4, 2, 4, 3), seed=9, mask_kind="bernoulli", zero_u=False, c_std=0.3)
    pb, g = _problem(case, dev)
    ks = KalmanStep(pb, g["eps"], use_graphs=False)
    t = ks.step()
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    assert abs(ret[0]["elbo"] - float(t[5])) <= 2e-6 * abs(float(t[5]))
    assert abs(ret[0]["elbo"] - ret[1]["elbo"]) == 0.0
    for k in ("dA", "dC"):
        assert rel(ret[0][k], ks.grads[k].cpu()) < 2e-5, k
        assert torch.equal(ret[0][k], ret[1][k])
    full = dict(dY=ks.grads["dY"].cpu(), dalpha=ks.grads["dalpha"].cpu())
    for r in (0, 1):
        lo, hi = shard_bounds(301, r, 2)
        for k in ("dY", "dalpha"):
            assert rel(ret[r][k], full[k][lo:hi]) < 2e-5, (r, k)


@pytest.mark.parametrize("switching", [False, True])
def test_impute_pipeline_matches_the_monolithic_call(switching):
    """engine.ImputePipeline (host-resident inputs, chunks of the batch moving through copy / compute / drain streams)
    returns exactly what one KalmanFilter.impute_observations call on the whole batch returns; last chunk partial."""
    from kalman_vae_b200 import KalmanFilter
    from kalman_vae_b200.dyn_param import PrecomputedWeights
    from kalman_vae_b200.engine import ImputePipeline
    dev = torch.device("cuda:0")
    shape = Shape(300, 24, 4, 2, 4, 3, switching, switching)
    case = make_case(shape, seed=31, mask_kind="block", zero_u=False, c_std=0.3)
    dyn = PrecomputedWeights(case["A"], case["B"], case["C"], case["Q"] if switching else None, switching=switching)
    kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, case["mu0"], case["Sigma0"], dyn).to(dev)
    host = {k: case[k].float().contiguous().pin_memory() for k in ("Y", "U", "mask", "alpha")}
    with torch.no_grad():
        dyn.set_weights(host["alpha"].to(dev))
        want_i, want_f, _, _ = kf.impute_observations(host["Y"].to(dev), host["U"].to(dev), host["mask"].to(dev))
    pipe = ImputePipeline(kf, chunk=128, want_filtered=True, device=dev)
    for _ in range(2):        # second run re-uses the slots
        got_i, got_f = pipe.run(host["Y"], host["mask"], alpha=host["alpha"], U=host["U"])
        assert torch.equal(got_i, want_i.cpu()) and torch.equal(got_f, want_f.cpu())
    assert bool((case["mask"] == 0).any()) and bool(torch.isfinite(got_i).all())
