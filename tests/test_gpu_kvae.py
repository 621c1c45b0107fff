"""GPU: the drop-in inside the REAL model.  The reference's KVAE module (kvae/model/model.py, unmodified, from
baseline/_ref or /root/reference) is built twice with identical weights -- once with its own KalmanFilter /
DynamicsParameter (stock ATen ops on the GPU) and once with the three names of INTEGRATION.md section 2 pointing at
kalman_vae_b200 -- and the training-step body of kvae/train/train.py:32-58 runs on both: same loss, same gradients on every
parameter (encoder, decoder, LSTM, A/B/C), same loss trajectory under Adam."""
import pytest
import torch

from kalman_vae_b200 import kvae_step

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(kvae_step.reference_root() is None, reason="reference sources not present (oracle/install_reference.sh)")]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _pair(dynamics):
    dev = torch.device("cuda:0")
    ref = kvae_step.ReferenceTrainStep(dev, drop_in=False, dynamics_model=dynamics, batch=32, T=20, seed=3)
    new = kvae_step.ReferenceTrainStep(dev, drop_in=True, dynamics_model=dynamics, batch=32, T=20, seed=3)
    missing = new.model.load_state_dict(ref.model.state_dict(), strict=True)   # interchangeable state dicts
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref, new


def test_state_dict_keys_identical():
    for dyn in ("lstm", "switching"):
        ref, new = _pair(dyn)
        assert list(ref.model.state_dict().keys()) == list(new.model.state_dict().keys())
        assert type(new.model.kalman_filter).__module__.startswith("kalman_vae_b200")


def test_lstm_kvae_loss_and_all_gradients_match_reference_on_gpu():
    ref, new = _pair("lstm")
    x = ref.synthetic_batch(seed=5).cuda()
    out = {}
    for name, st in (("ref", ref), ("new", new)):
        torch.manual_seed(1234)          # same encoder noise and the same rsample draw (kalman_filter.py:351) in both
        m = st.model
        m.train()
        m.kalman_filter.dyn_params.reset_state()
        mask = torch.ones(32, 20, device=x.device)
        m.zero_grad(set_to_none=True)
        o = m(x, mask=mask)
        losses = m.compute_loss(x, o, mask=mask)
        losses["loss"].backward()
        out[name] = (losses, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}, o)
    lr, gr, orf = out["ref"]
    ln, gn, onw = out["new"]
    assert rel(ln["elbo_kf"], lr["elbo_kf"]) < 2e-5, (float(ln["elbo_kf"]), float(lr["elbo_kf"]))
    assert rel(ln["loss"], lr["loss"]) < 2e-5
    for k in ("mus_smooth", "Sigmas_smooth", "mus_filt", "Sigmas_filt"):
        assert rel(onw[k], orf[k]) < 5e-5, k
    assert set(gr) == set(gn)
    worst = 0.0
    for k in gr:
        if float(gr[k].abs().max()) == 0.0:
            assert float(gn[k].abs().max()) < 1e-6, k
            continue
        e = rel(gn[k], gr[k])
        worst = max(worst, e)
        assert e < 5e-4, (k, e)     # both sides are fp32 (the reference side through cuDNN + ~10^4 ATen ops)
    print(f"KVAE (lstm) drop-in vs reference ops on the GPU: loss rel {rel(ln['loss'], lr['loss']):.1e}, worst gradient rel {worst:.1e} over {len(gr)} parameters")


def test_three_adam_steps_follow_the_reference_trajectory():
    ref, new = _pair("lstm")
    xs = [ref.synthetic_batch(seed=10 + i).cuda() for i in range(3)]
    lr, ln = [], []
    for st, acc in ((ref, lr), (new, ln)):
        torch.manual_seed(77)
        for x in xs:
            acc.append(float(st.step(x)))
    for a, b in zip(ln, lr):
        assert abs(a - b) <= 2e-3 * abs(b), (ln, lr)
    print("loss trajectories:", ln, lr)


def test_switching_kvae_trains_with_the_drop_in():
    """SKVAE: the regime chain draws its Gumbel noise in one call here and per step in the reference, so the two runs are
    different samples of the same model: check that the step runs, the loss is finite and decreases on a fixed batch."""
    _, new = _pair("switching")
    x = new.synthetic_batch(seed=2).cuda()
    torch.manual_seed(5)
    losses = [float(new.step(x)) for _ in range(8)]
    assert all(l == l and abs(l) < 1e9 for l in losses), losses
    assert min(losses[4:]) < losses[0], losses


def test_switching_reference_rng_samples_the_same_regimes():
    """SwitchingDynamicsParameter(reference_rng=True): the Gumbel noise is drawn per step in the reference's order
    (switch_dyn_param.py:52,69), so a drop-in run seeded like a reference-on-CUDA run picks IDENTICAL regimes (eval mode:
    hard one-hot samples) and the same smoothed states."""
    ref, new = _pair("switching")
    new.model.kalman_filter.dyn_params.reference_rng = True
    x = ref.synthetic_batch(seed=9).cuda()
    outs = {}
    for name, st in (("ref", ref), ("new", new)):
        m = st.model
        m.eval()
        with torch.no_grad():
            torch.manual_seed(4321)
            m.kalman_filter.dyn_params.reset_state()
            outs[name] = m(x, mask=torch.ones(32, 20, device=x.device))
    yr, yn = outs["ref"]["state_probs"], outs["new"]["state_probs"]
    assert torch.equal(yr.argmax(-1), yn.argmax(-1))                      # identical regime picks
    assert float((yr - yn).abs().max()) < 1e-5
    assert rel(outs["new"]["mus_smooth"], outs["ref"]["mus_smooth"]) < 5e-5
    assert rel(outs["new"]["Sigmas_smooth"], outs["ref"]["Sigmas_smooth"]) < 5e-5


def test_graphed_train_step_follows_the_eager_trajectory():
    """kvae_step.GraphedTrainStep (forward + loss + backward + clip + Adam captured in one CUDA graph, replayed) against
    the eager step body of train.py:32-58 on the same model: same weights, same seed, same batches -> same losses."""
    dev = torch.device("cuda:0")
    eager = kvae_step.ReferenceTrainStep(dev, drop_in=True, dynamics_model="lstm", batch=32, T=20, seed=3)
    graphed = kvae_step.GraphedTrainStep(dev, dynamics_model="lstm", batch=32, T=20, seed=3)
    init = {k: v.clone() for k, v in eager.model.state_dict().items()}
    xs = [eager.synthetic_batch(seed=10 + i).to(dev) for i in range(4)]
    graphed.capture(xs[0])                      # the warm-up inside moved the weights and the Adam state: rewind both
    graphed.model.load_state_dict(init, strict=True)
    for st in graphed.opt.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    le, lg = [], []
    torch.manual_seed(77)
    for x in xs:
        le.append(float(eager.step(x)))
    torch.manual_seed(77)
    for x in xs:
        lg.append(float(graphed.step(x)))
    print("loss trajectories (graphed, eager):", lg, le)
    for a, b in zip(lg, le):
        assert abs(a - b) <= 2e-3 * abs(b), (lg, le)
    assert int(graphed.model.kalman_filter.dyn_params.A.grad is not None)


def test_graphed_train_step_switching_dynamics():
    """The SKVAE (regime sampler kernels + bi-GRU posterior) through the same capture: the step replays, the loss is
    finite and goes down on a fixed batch, and no factorisation failed (status word)."""
    dev = torch.device("cuda:0")
    st = kvae_step.GraphedTrainStep(dev, dynamics_model="switching", batch=32, T=20, seed=3)
    x = st.synthetic_batch(seed=2).to(dev)
    st.capture(x)
    torch.manual_seed(5)
    losses = [float(st.step(x)) for _ in range(8)]
    st.check()
    assert all(l == l and abs(l) < 1e9 for l in losses), losses
    assert min(losses[4:]) < losses[0], losses


def test_lstm_dynamics_with_missing_observations_under_autograd_matches_reference():
    """kf.strict = True: lstm dynamics + a mask with zeros + gradients wanted takes the step-by-step autograd path
    (general_steps.py; the gradient passes through the LSTM between the steps).  Against the reference's own KalmanFilter
    + DynamicsParameter on the same GPU with the same weights: the nine outputs, the ELBO and the gradients w.r.t. the
    observations, A, B, C and the LSTM / head weights."""
    import kalman_vae_b200
    from kalman_vae_b200.synthetic import Shape, make_case
    model_mod, _ = kvae_step.load_reference_model_module()
    dev = torch.device("cuda:0")
    case = make_case(Shape(24, 12, 4, 2, 4, 3), seed=4, mask_kind="block", zero_u=False, c_std=0.3)
    assert bool((case["mask"] == 0).any())
    g = lambda k: case[k].to(dev).float()
    torch.manual_seed(11)
    ref_dyn = model_mod.base_dyn_param.DynamicsParameter(g("A"), g("B"), g("C")).to(dev)
    new_dyn = kalman_vae_b200.DynamicsParameter(g("A"), g("B"), g("C")).to(dev)
    with torch.no_grad():   # a head that actually mixes the modes
        ref_dyn.head_w.bias.copy_(torch.tensor([0.0, -0.5, 0.3], device=dev))
    new_dyn.load_state_dict(ref_dyn.state_dict(), strict=True)
    ref_kf = model_mod.KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, g("mu0"), g("Sigma0"), ref_dyn).to(dev)
    new_kf = kalman_vae_b200.KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, g("mu0"), g("Sigma0"), new_dyn).to(dev)
    new_kf.strict = True
    eps = torch.randn(24, 12, 4, device=dev)
    res = {}
    for tag, kf, dyn in (("ref", ref_kf, ref_dyn), ("new", new_kf, new_dyn)):
        Y = g("Y").requires_grad_(True)
        dyn.reset_state()
        outs = kf.smooth(Y, g("U"), g("mask"))
        if tag == "new":
            kf._draw_eps = lambda B, T, n, like: eps
            val = kf.elbo(outs[0], outs[1], Y, g("U"), outs[6], outs[7], outs[8], mask=g("mask"))
        else:   # the reference draws inside MultivariateNormal.rsample: inject the same standard-normal draw
            import torch.distributions as D
            orig = D.MultivariateNormal.rsample
            D.MultivariateNormal.rsample = lambda self, sample_shape=torch.Size(): self.loc + (self._unbroadcasted_scale_tril @ eps.unsqueeze(-1)).squeeze(-1)
            try:
                val = kf.elbo(outs[0], outs[1], Y, g("U"), outs[6], outs[7], outs[8], mask=g("mask"))
            finally:
                D.MultivariateNormal.rsample = orig
        params = [Y, dyn.A, dyn.B, dyn.C, dyn.lstm.weight_ih_l0, dyn.lstm.weight_hh_l0, dyn.head_w.weight]
        res[tag] = ([o.detach() for o in outs], val.detach(), torch.autograd.grad(val, params))
    for a, b in zip(res["new"][0], res["ref"][0]):
        assert rel(a, b) < 2e-5, rel(a, b)
    assert abs(float(res["new"][1]) - float(res["ref"][1])) <= 2e-5 * abs(float(res["ref"][1]))
    for nm, a, b in zip(("dY", "dA", "dB", "dC", "dW_ih", "dW_hh", "dW_head"), res["new"][2], res["ref"][2]):
        print(f"lstm + mask under autograd, {nm}: rel {rel(a, b):.1e}")
        assert rel(a, b) < 2e-4, (nm, rel(a, b))      # measured: 1e-7 .. 2e-5
