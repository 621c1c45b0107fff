"""GPU parity tests proper: the CUDA path through the C ABI against the reference's golden vectors
(fp32 and fp64 runs of the unmodified reference), every instantiated lane count."""
import pytest
import torch

from kalman_vae_b200 import capi
from kalman_vae_b200 import functional as F
from kalman_vae_b200.functional import Problem
from tests._util import GRAD_NAMES, OUT_NAMES, check_close, golden_names, load_golden

pytestmark = pytest.mark.gpu


def lanes_for(n):
    return {2: (1, 2), 4: (1, 2, 4), 8: (4, 8), 16: (8, 16)}[n]


def problem(case, lanes, dev):
    g = {k: (v.to(dev).float().contiguous() if torch.is_tensor(v) else v) for k, v in case.items()}
    return Problem(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"],
                   bool(case["q_per_mode"]), bool(case["c_shared"]), lanes=lanes), g


def all_cases():
    out = []
    for name in golden_names():
        case = load_golden(name)[0]
        for lanes in lanes_for(case["A"].shape[-1]):
            out.append((name, lanes))
    return out


@pytest.mark.parametrize("name,lanes", all_cases())
def test_forward_elbo_backward_match_reference(name, lanes):
    dev = torch.device("cuda:0")
    case, cot, r32, r64 = load_golden(name)
    pb, g = problem(case, lanes, dev)
    F.info_word(dev).zero_()
    st, A_list, B_list, C_list = F.smooth_fwd(pb)
    got = dict(mus_smooth=st.mus_smooth, Sigmas_smooth=st.Sigmas_smooth, mus_filt=st.mus_filt, Sigmas_filt=st.Sigmas_filt,
               mus_pred=st.mus_pred, Sigmas_pred=st.Sigmas_pred, A_list=A_list, B_list=B_list, C_list=C_list)
    for k in OUT_NAMES:
        check_close(f"{name}.L{lanes}.{k}", got[k], r32[k], r64[k])
    if "elbo" not in r32 and "dY" not in r32:
        return
    terms = gel = None
    if "elbo" in r32:
        terms = F.elbo_terms(pb, st, g["eps"])
        check_close(f"{name}.L{lanes}.elbo", terms[5], r32["elbo"], r64["elbo"])
        gel = torch.ones(1, device=dev)
    cot_d = {k: v.to(dev).float().contiguous() for k, v in cot.items()} if cot else None
    gr = F.adjoint(pb, st, eps=g["eps"], g_elbo=gel, terms=terms, cot=cot_d)
    torch.cuda.synchronize()
    assert int(F.info_word(dev)) == 0
    gr = dict(dY=gr["dY"], dU=gr["dU"], dalpha=gr["dalpha"], dA=gr["dA"], dB=gr["dBm"], dC=gr["dC"], dQ=gr["dQ"])
    for k in GRAD_NAMES:
        if k in r32:
            check_close(f"{name}.L{lanes}.{k}", gr[k], r32[k], r64[k])
    if "elbo" in r32:
        # fused value + adjoint launch (KVAE_FLAG_WITH_ELBO): the ELBO against the reference, and the ELBO-only gradients
        # against the two-call sequence (the goldens' gradients include the dense cotangents, which the fused mode excludes)
        t_f = torch.empty(8, device=dev)
        fused = F.adjoint(pb, st, eps=g["eps"], g_elbo=gel, terms=t_f, with_elbo=True)
        plain = F.adjoint(pb, st, eps=g["eps"], g_elbo=gel, terms=terms)
        torch.cuda.synchronize()
        check_close(f"{name}.L{lanes}.elbo_fused", t_f[5], r32["elbo"], r64["elbo"])
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
        for k in ("dY", "dU", "dalpha", "dA", "dBm", "dC", "dQ"):
            if fused[k] is not None:
                assert rel(fused[k], plain[k]) < 5e-5, (name, lanes, k)   # fp32 noise floor of the cancelling sums


@pytest.mark.parametrize("lanes", [1, 2, 4])
def test_mask_zero_bit_exact(lanes):
    """mask handling is bit-exact: mask=0 => mu_filt == mu_pred, Sigma_filt == (Sigma_pred + Sigma_pred^T)/2."""
    dev = torch.device("cuda:0")
    case = load_golden("kalman_zero_mask")[0]
    pb, _ = problem(case, lanes, dev)
    st, *_ = F.smooth_fwd(pb)
    assert torch.equal(st.mus_filt, st.mus_pred)
    assert torch.equal(st.Sigmas_filt, 0.5 * (st.Sigmas_pred + st.Sigmas_pred.mT))


@pytest.mark.parametrize("lanes", [1, 4])
def test_ragged_batch_tail_and_filter_only(lanes):
    """B not a multiple of the sequences per CTA; filter-only call leaves the smoothed buffers untouched."""
    from kalman_vae_b200.synthetic import Shape, make_case
    from oracle import kalman_oracle as ko
    dev = torch.device("cuda:0")
    shape = Shape(131, 6, 4, 2, 4, 3)
    case = make_case(shape, seed=4, mask_kind="bernoulli", zero_u=False, c_std=0.3)
    pb, _ = problem(case, lanes, dev)
    st, *_ = F.smooth_fwd(pb, smooth=False)
    assert st.mus_smooth is None
    r32 = ko.run_case(case, torch.float32, want_grads=False)
    r64 = ko.run_case(case, torch.float64, want_grads=False)
    for k, v in (("mus_filt", st.mus_filt), ("Sigmas_filt", st.Sigmas_filt), ("mus_pred", st.mus_pred), ("Sigmas_pred", st.Sigmas_pred)):
        check_close(k, v, r32[k], r64[k])


def test_full_size_properties_cfg2():
    """BASELINE cfg2 size (B=8192,T=20): size-independent properties instead of an oracle run:
    symmetry of the filtered/smoothed covariances, lane-count invariance, mask=1 innovation identity
    (smoothed == filtered at T-1), and linearity of the adjoint in the upstream gradient."""
    from kalman_vae_b200.synthetic import CONFIGS, make_case
    dev = torch.device("cuda:0")
    shape = CONFIGS["cfg2"]
    case = make_case(shape, seed=10)
    res = {}
    for lanes in (1, 4):
        pb, g = problem(case, lanes, dev)
        st, *_ = F.smooth_fwd(pb)
        assert torch.equal(st.Sigmas_filt, st.Sigmas_filt.mT.contiguous())
        assert torch.equal(st.Sigmas_smooth[:, :-1], st.Sigmas_smooth[:, :-1].mT.contiguous())
        assert torch.equal(st.mus_smooth[:, -1], st.mus_filt[:, -1])
        terms = F.elbo_terms(pb, st, g["eps"])
        g1 = F.adjoint(pb, st, eps=g["eps"], g_elbo=torch.ones(1, device=dev), terms=terms)
        g2 = F.adjoint(pb, st, eps=g["eps"], g_elbo=torch.full((1,), 2.0, device=dev), terms=terms)
        for k in ("dY", "dalpha", "dA", "dC"):
            assert torch.allclose(2.0 * g1[k], g2[k], rtol=1e-5, atol=1e-7), k
        res[lanes] = (st, terms, g1)
    # lane count changes only the order of a few reductions
    for k in ("mus_smooth", "Sigmas_smooth"):
        a, b = getattr(res[1][0], k), getattr(res[4][0], k)
        assert float((a - b).norm() / b.norm()) < 1e-5, k
    assert abs(float(res[1][1][5]) - float(res[4][1][5])) <= 1e-5 * abs(float(res[4][1][5]))
    for k in ("dY", "dalpha", "dA"):
        a, b = res[1][2][k], res[4][2][k]
        assert float((a - b).norm() / b.norm()) < 1e-4, k


def _elbo_only_case(name):
    case = load_golden(name)[0]
    return case


@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_lstm_default", "kalman_switch", "kalman_fractional", "kalman_zero_mask"])
@pytest.mark.parametrize("lanes", [1, 4])
def test_training_gradients_match_oracle(name, lanes):
    """The training case proper: gradient of the ELBO alone (no cotangents on the smooth outputs) in the fused
    value + adjoint launch.  With lanes = 1 and T % 4 == 0 this is the thread-per-sequence kernel pair
    (csrc/kvae_seq.cuh, kvae_seq_bwd.cuh); the oracle (same op order as the reference, pinned by tests/test_oracle.py)
    supplies fp32 and fp64 answers for exactly this loss."""
    from oracle import kalman_oracle as ko
    dev = torch.device("cuda:0")
    case = _elbo_only_case(name)
    r32 = ko.run_case(case, torch.float32, want_grads=True)
    r64 = ko.run_case(case, torch.float64, want_grads=True)
    pb, g = problem(case, lanes, dev)
    F.info_word(dev).zero_()
    st, *_ = F.smooth_fwd(pb)
    t_f = torch.empty(8, device=dev)
    gr = F.adjoint(pb, st, eps=g["eps"], g_elbo=torch.ones(1, device=dev), terms=t_f, with_elbo=True)
    torch.cuda.synchronize()
    assert int(F.info_word(dev)) == 0
    check_close(f"{name}.L{lanes}.elbo", t_f[5], r32["elbo"], r64["elbo"])
    gr = dict(dY=gr["dY"], dU=gr["dU"], dalpha=gr["dalpha"], dA=gr["dA"], dB=gr["dBm"], dC=gr["dC"], dQ=gr["dQ"])
    for k in GRAD_NAMES:
        if k in r32 and gr[k] is not None and float(r64[k].abs().max()) > 0:
            e32, e64, floor = check_close(f"{name}.L{lanes}.{k}", gr[k], r32[k], r64[k])
            print(f"{name}.L{lanes}.{k}: e32 {e32:.1e} e64 {e64:.1e} floor {floor:.1e}")


def test_full_size_cfg2_against_oracle():
    """BASELINE configs[1] at FULL size (B=8192, T=20): all nine outputs, the ELBO and the training gradients against
    the CPU oracle in fp32 and fp64 (the oracle needs ~1 s for this size), default lane count."""
    from kalman_vae_b200.synthetic import CONFIGS, make_case
    from oracle import kalman_oracle as ko
    dev = torch.device("cuda:0")
    case = make_case(CONFIGS["cfg2"], seed=10)
    r32 = ko.run_case(case, torch.float32, want_grads=True)
    r64 = ko.run_case(case, torch.float64, want_grads=True)
    pb, g = problem(case, 0, dev)
    F.info_word(dev).zero_()
    st, A_list, B_list, C_list = F.smooth_fwd(pb)
    got = dict(mus_smooth=st.mus_smooth, Sigmas_smooth=st.Sigmas_smooth, mus_filt=st.mus_filt, Sigmas_filt=st.Sigmas_filt,
               mus_pred=st.mus_pred, Sigmas_pred=st.Sigmas_pred, A_list=A_list, B_list=B_list, C_list=C_list)
    worst = 0.0
    for k in OUT_NAMES:
        e32, e64, floor = check_close(f"cfg2.{k}", got[k], r32[k], r64[k])
        worst = max(worst, e32)
    t_f = torch.empty(8, device=dev)
    gr = F.adjoint(pb, st, eps=g["eps"], g_elbo=torch.ones(1, device=dev), terms=t_f, with_elbo=True)
    torch.cuda.synchronize()
    assert int(F.info_word(dev)) == 0
    check_close("cfg2.elbo", t_f[5], r32["elbo"], r64["elbo"])
    gr = dict(dY=gr["dY"], dalpha=gr["dalpha"], dA=gr["dA"], dB=gr["dBm"], dC=gr["dC"])
    for k, v in gr.items():
        if float(r64[k].abs().max()) > 0:
            e32, e64, floor = check_close(f"cfg2.{k}", v, r32[k], r64[k])
            print(f"cfg2 full size {k}: e32 {e32:.1e} e64 {e64:.1e} floor {floor:.1e}")
    print(f"cfg2 full size: worst forward error vs reference-order fp32 {worst:.1e}")


@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_switch", "kalman_zero_mask", "kalman_T1", "kalman_n8"])
def test_projections_emitted_by_the_sweeps(name):
    """a_filt = C_t mu_{t|t} and a_smooth = C_t mu_{t|T} written by the filter / smoother sweeps
    (kvae_states.a_filt / a_smooth) equal the reference's `(C_list @ mus).squeeze(-1)` (model.py:280-281, 287-288),
    every lane count (lanes = 1 with T % 4 == 0: the thread-per-sequence kernel)."""
    dev = torch.device("cuda:0")
    case, _, r32, r64 = load_golden(name)
    want = {k: (r[("C_list")].double() @ r[m].double()).squeeze(-1) for k, m in (("a_filt", "mus_filt"), ("a_smooth", "mus_smooth"))
            for r in (r64,)}
    want32 = {k: (r32["C_list"] @ r32[m]).squeeze(-1) for k, m in (("a_filt", "mus_filt"), ("a_smooth", "mus_smooth"))}
    for lanes in lanes_for(case["A"].shape[-1]):
        pb, _ = problem(case, lanes, dev)
        st, *_ = F.smooth_fwd(pb, smooth=True, lists=False, projections=True)
        for k in ("a_filt", "a_smooth"):
            check_close(f"{name}.L{lanes}.{k}", getattr(st, k), want32[k], want[k], rtol=2e-5)


# shapes outside kvae_configs.h: built on demand from the same sources (kalman_vae_b200/build.py build_shape_lib;
# __graft_entry__.build() pre-builds exactly these so that the GPU box only loads them)
ON_DEMAND_SHAPES = [(3, 2, 3, 2), (6, 3, 5, 2)]


@pytest.mark.parametrize("dims", ON_DEMAND_SHAPES)
@pytest.mark.parametrize("switching", [False, True])
def test_shape_built_on_demand_matches_oracle(dims, switching):
    """KVAEConfig allows any (a_dim, z_dim, u_dim, num_modes) (kvae/utils/config.py:4-60).  A tuple that the default
    library does not instantiate -- here z_dim = 3 (odd: one lane per sequence) and z_dim = 6 (two lanes) -- gets its own
    build of the same kernels; forward outputs, ELBO and training gradients against the CPU oracle in fp32 / fp64."""
    from kalman_vae_b200 import capi
    from kalman_vae_b200.synthetic import Shape, make_case
    from oracle import kalman_oracle as ko
    n, p, m, K = dims
    dev = torch.device("cuda:0")
    assert not capi.lib().kvae_supported(capi.make_dims(1, 1, n, p, m, K, switching, switching)), "shape is in the default library"
    case = make_case(Shape(70, 11, n, p, m, K, switching, switching), seed=5, mask_kind="bernoulli", zero_u=False, c_std=0.3)
    r32 = ko.run_case(case, torch.float32, want_grads=True)
    r64 = ko.run_case(case, torch.float64, want_grads=True)
    pb, g = problem(case, 0, dev)
    assert capi.lib_for(pb.dims) is not capi.lib()
    F.info_word(dev).zero_()
    st, A_list, B_list, C_list = F.smooth_fwd(pb)
    got = dict(mus_smooth=st.mus_smooth, Sigmas_smooth=st.Sigmas_smooth, mus_filt=st.mus_filt, Sigmas_filt=st.Sigmas_filt,
               mus_pred=st.mus_pred, Sigmas_pred=st.Sigmas_pred, A_list=A_list, B_list=B_list, C_list=C_list)
    for k in OUT_NAMES:
        check_close(f"{dims}.{k}", got[k], r32[k], r64[k])
    t_f = torch.empty(8, device=dev)
    pb.mask_partials = st.mask_partials
    gr = F.adjoint(pb, st, eps=g["eps"], g_elbo=torch.ones(1, device=dev), terms=t_f, with_elbo=True)
    torch.cuda.synchronize()
    assert int(F.info_word(dev)) == 0
    check_close(f"{dims}.elbo", t_f[5], r32["elbo"], r64["elbo"])
    gr = dict(dY=gr["dY"], dU=gr["dU"], dalpha=gr["dalpha"], dA=gr["dA"], dB=gr["dBm"], dC=gr["dC"], dQ=gr["dQ"])
    for k, v in gr.items():
        if v is not None and k in r64 and float(r64[k].abs().max()) > 0:
            e32, e64, floor = check_close(f"{dims}.{k}", v, r32[k], r64[k])
            print(f"{dims} switching={switching} {k}: e32 {e32:.1e} e64 {e64:.1e} floor {floor:.1e}")
