"""CPU test: the C-ABI shared library loads and exports every symbol include/kvae_kalman.h declares
(no compute calls without a GPU), and the host-side argument checks reject bad shapes."""
import ctypes
import os
import re

import pytest

from kalman_vae_b200 import build as kbuild
from kalman_vae_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    kbuild.build()
    return ctypes.CDLL(capi.LIB_PATH)


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "kvae_kalman.h")).read()
    declared = set(re.findall(r"\b(kvae_[a-z_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_abi_version_and_support_table(lib):
    L = capi.lib()
    assert L.kvae_abi_version() == 6
    ok = capi.make_dims(8, 20, 4, 2, 4, 3, False, False, 0)
    assert capi.supported(ok)
    assert capi.pick_lanes(ok) == 4                      # small batch -> widest lane group
    assert capi.pick_lanes(capi.make_dims(65536, 1000, 4, 2, 4, 3, False, False, 0)) == 1   # >= 12288 sequences, T % 4 == 0: thread per sequence
    assert capi.pick_lanes(capi.make_dims(65536, 1001, 4, 2, 4, 3, False, False, 0)) == 4   # T % 4 != 0: lane groups
    assert capi.supported(capi.make_dims(8, 20, 16, 8, 16, 8, True, True, 16))
    raw = lambda d: bool(L.kvae_supported(ctypes.byref(d)))
    assert not raw(capi.make_dims(8, 20, 5, 2, 4, 3, False, False, 0))   # not in the default library (capi.supported would build it on demand)
    assert not capi.supported(capi.make_dims(8, 20, 4, 2, 4, 3, True, False, 0))      # mixed variant
    assert not capi.supported(capi.make_dims(8, 20, 4, 2, 4, 3, False, False, 3))     # lanes must divide n


def test_null_arguments_are_rejected_without_touching_the_gpu(lib):
    L = capi.lib()
    rc = L.kvae_kf_filter_smooth_fwd(None, None, None, None, None, None, None, 0, None)
    assert rc < 0 and b"null" in L.kvae_last_error()


def test_cpu_tensors_raise():
    import torch
    from kalman_vae_b200.functional import Problem
    from kalman_vae_b200.synthetic import Shape, make_case
    c = make_case(Shape(2, 3, 4, 2, 4, 3))
    pb = Problem(c["Y"], c["U"], c["mask"], c["alpha"], c["A"], c["B"], c["C"], c["Q"], c["R"], c["mu0"], c["Sigma0"], False, False)
    with pytest.raises(capi.KvaeError):
        pb.inputs()        # CPU tensors: no CPU implementation exists


def test_data_parallel_and_regime_entries_reject_bad_arguments(lib):
    """kvae_dp_* / kvae_kf_bwd_dp / kvae_regime_*: argument checks run before anything touches a device."""
    L = capi.lib()
    assert L.kvae_dp_handle_bytes() == 64                                   # sizeof(cudaIpcMemHandle_t)
    assert L.kvae_dp_create(0, 0, 0, 10, None, None) < 0                    # null out pointers / world < 1
    assert b"kvae_dp_create" in L.kvae_dp_last_error()
    assert L.kvae_dp_connect(None, None) < 0
    assert L.kvae_dp_finalize(None, None, None, None, None, None) < 0
    rc = L.kvae_kf_bwd_dp(None, None, None, None, ctypes.c_float(1e-6), None, None, None, None, None, 0, None, None)
    assert rc < 0 and b"communicator" in L.kvae_last_error()
    assert L.kvae_kf_mask_partials_count(None) == 0
    d = capi.make_dims(8192, 20, 4, 2, 4, 3, False, False, 0)
    assert capi.mask_partials_count(d) == 8192 // (128 // 4)                # one partial per forward CTA (128 threads, L = 4)
    assert L.kvae_kf_mask_partials_count(ctypes.byref(capi.make_dims(8, 20, 5, 2, 4, 3, False, False, 0))) == 0   # shape not in the default library


def test_shape_on_demand_is_built_and_loaded(lib):
    """A (n, p, m, K) tuple outside kvae_configs.h: capi.lib_for compiles the same sources for it (nvcc, cached under
    kalman_vae_b200/_jit/) and the resulting library exports the whole ABI; KVAE_JIT=0 keeps the old behaviour."""
    d = capi.make_dims(8, 20, 3, 2, 3, 2, False, False, 0)
    assert not capi.lib().kvae_supported(ctypes.byref(d))
    L = capi.lib_for(d)
    assert L is not capi.lib()
    assert capi.supported(d) and capi.supported(capi.make_dims(8, 20, 3, 2, 3, 2, True, True, 0))
    assert capi.pick_lanes(d) == 1                                            # odd z_dim: one lane per sequence
    assert not capi.supported(capi.make_dims(8, 20, 3, 2, 3, 2, False, False, 2))
    assert capi.mask_partials_count(d) > 0 and capi.bwd_workspace_bytes(d) > 0
    for s in capi.EXPORTED_SYMBOLS:
        assert hasattr(L, s), s
