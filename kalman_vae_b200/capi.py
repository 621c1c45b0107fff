"""ctypes binding of libkvae_kalman.so (include/kvae_kalman.h) for torch CUDA tensors.

This is the only place the package touches the native library.  There is no fallback: if the
library is missing or the tensors are not CUDA tensors the call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KVAE_LIB") or os.path.join(_HERE, "libkvae_kalman.so")   # KVAE_LIB: development A/B builds


class KvaeDims(Structure):
    _fields_ = [(k, c_int32) for k in ("B", "T", "n", "p", "m", "K", "q_per_mode", "c_shared", "lanes", "flags")]


class KvaeInputs(Structure):
    _fields_ = [(k, c_void_p) for k in ("Y", "U", "mask", "alpha", "A", "Bm", "C", "Q", "R", "mu0", "Sigma0",
                                        "mu_init", "Sigma_init", "A_dense", "B_dense", "C_dense", "Q_dense")]


class KvaeStates(Structure):
    _fields_ = [(k, c_void_p) for k in ("mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred",
                                        "mus_smooth", "Sigmas_smooth", "mask_partials", "a_filt", "a_smooth")]


class KvaeCotangents(Structure):
    _fields_ = [(k, c_void_p) for k in ("mus_smooth", "Sigmas_smooth", "mus_filt", "Sigmas_filt",
                                        "mus_pred", "Sigmas_pred", "A_list", "B_list", "C_list")]


class KvaeGrads(Structure):
    _fields_ = [(k, c_void_p) for k in ("dY", "dU", "dalpha", "dA", "dBm", "dC", "dQ", "dmus", "dSigmas")]


class KvaeRegimeDims(Structure):
    _fields_ = [("B", c_int32), ("T", c_int32), ("K", c_int32), ("hard", c_int32), ("tau", c_float)]


class KvaeLstm(Structure):
    _fields_ = [(k, c_void_p) for k in ("w_ih", "w_hh", "b_ih", "b_hh", "w_head", "b_head", "h0", "c0", "h_out", "c_out")] + \
               [("hidden", c_int32)]


class KvaeCholOpts(Structure):
    _fields_ = [("jitter_q", c_float), ("diag_smooth", c_int32), ("diag_q", c_int32)]


INFO_PIVOT, INFO_PEER, INFO_CHOL_S, INFO_CHOL_Q = 1, 2, 4, 8


class KvaeError(RuntimeError):
    pass


_lib = None


def _load(path):
    """dlopen + ctypes signatures of one build of the library (the default one or a shape built on demand)."""
    L = ctypes.CDLL(path)
    L.kvae_abi_version.restype = c_int
    L.kvae_last_error.restype = c_char_p
    L.kvae_supported.argtypes = [POINTER(KvaeDims)]
    L.kvae_pick_lanes.argtypes = [POINTER(KvaeDims)]
    L.kvae_kf_filter_smooth_fwd.argtypes = [POINTER(KvaeDims), POINTER(KvaeInputs), POINTER(KvaeStates),
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    L.kvae_kf_filter_lstm_fwd.argtypes = [POINTER(KvaeDims), POINTER(KvaeInputs), POINTER(KvaeStates), c_void_p, c_void_p,
                                          c_void_p, POINTER(KvaeLstm), c_void_p, c_void_p, c_int, c_void_p]
    L.kvae_kf_elbo_workspace_bytes.argtypes = [POINTER(KvaeDims)]
    L.kvae_kf_elbo_workspace_bytes.restype = c_size_t
    L.kvae_kf_elbo_fwd.argtypes = [POINTER(KvaeDims), POINTER(KvaeInputs), POINTER(KvaeStates), c_void_p, c_float,
                                   c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    L.kvae_kf_elbo_fwd_ex.argtypes = [POINTER(KvaeDims), POINTER(KvaeInputs), POINTER(KvaeStates), c_void_p, c_float,
                                      POINTER(KvaeCholOpts), c_void_p, c_void_p, c_void_p, c_int, c_void_p]
    L.kvae_kf_bwd_ex.argtypes = [POINTER(KvaeDims), POINTER(KvaeInputs), POINTER(KvaeStates), c_void_p, c_float,
                                 POINTER(KvaeCholOpts), c_void_p, c_void_p, POINTER(KvaeCotangents), POINTER(KvaeGrads),
                                 c_void_p, c_void_p, c_int, c_void_p]
    L.kvae_kf_bwd_workspace_bytes.argtypes = [POINTER(KvaeDims)]
    L.kvae_kf_bwd_workspace_bytes.restype = c_size_t
    L.kvae_kf_bwd.argtypes = [POINTER(KvaeDims), POINTER(KvaeInputs), POINTER(KvaeStates), c_void_p, c_float,
                              c_void_p, c_void_p, POINTER(KvaeCotangents), POINTER(KvaeGrads), c_void_p,
                              c_void_p, c_int, c_void_p]
    L.kvae_dp_last_error.restype = c_char_p
    L.kvae_dp_handle_bytes.restype = c_size_t
    L.kvae_dp_create.argtypes = [c_int, c_int, c_int, c_size_t, POINTER(c_void_p), c_void_p]
    L.kvae_dp_connect.argtypes = [c_void_p, c_void_p]
    L.kvae_dp_destroy.argtypes = [c_void_p]
    L.kvae_dp_finalize.argtypes = [POINTER(KvaeDims), c_void_p, POINTER(KvaeGrads), c_void_p, c_void_p, c_void_p]
    L.kvae_kf_bwd_dp.argtypes = [POINTER(KvaeDims), POINTER(KvaeInputs), POINTER(KvaeStates), c_void_p, c_float, c_void_p,
                                 c_void_p, POINTER(KvaeGrads), c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    L.kvae_regime_last_error.restype = c_char_p
    L.kvae_regime_supported.argtypes = [c_int]
    L.kvae_regime_sample_fwd.argtypes = [POINTER(KvaeRegimeDims)] + [c_void_p] * 7 + [c_int, c_void_p]
    L.kvae_regime_sample_bwd.argtypes = [POINTER(KvaeRegimeDims)] + [c_void_p] * 10 + [c_int, c_void_p]
    L.kvae_kf_mask_partials_count.argtypes = [POINTER(KvaeDims)]
    L.kvae_kf_mask_partials_count.restype = c_size_t
    if L.kvae_abi_version() != 6:
        raise KvaeError(f"{path}: ABI version mismatch")
    return L


def lib():
    """Loads the native library (raises if it has not been built: there is no CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KvaeError(
            f"{LIB_PATH} not found: build it with `python -m kalman_vae_b200.build` "
            "(or __graft_entry__.build()). The Kalman hot path has no CPU fallback.")
    _lib = _load(LIB_PATH)
    return _lib


# Shapes on demand: (n, p, m, K) tuples outside kvae_configs.h get their own build of the same sources
# (kalman_vae_b200/build.py build_shape_lib; KVAE_JIT=0 switches this off and such shapes raise).
_shape_libs = {}


def lib_for(dims):
    """The library build that instantiates dims' (n, p, m, K): the default one, or one compiled on demand."""
    key = (dims.n, dims.p, dims.m, dims.K)
    L = _shape_libs.get(key)
    if L is not None:
        return L
    L = lib()
    probe = KvaeDims(1, 1, dims.n, dims.p, dims.m, dims.K, 0, 0, 0, 0)
    if not L.kvae_supported(byref(probe)) and os.environ.get("KVAE_JIT", "1") != "0" and min(key) >= 1 and dims.n <= 16:
        from . import build as kbuild
        L = _load(kbuild.build_shape_lib(*key))
    _shape_libs[key] = L
    return L


EXPORTED_SYMBOLS = [
    "kvae_abi_version", "kvae_last_error", "kvae_supported", "kvae_pick_lanes",
    "kvae_kf_mask_partials_count", "kvae_kf_filter_smooth_fwd", "kvae_kf_filter_lstm_fwd", "kvae_kf_elbo_workspace_bytes", "kvae_kf_elbo_fwd",
    "kvae_kf_bwd_workspace_bytes", "kvae_kf_bwd", "kvae_kf_elbo_fwd_ex", "kvae_kf_bwd_ex",
    "kvae_regime_last_error", "kvae_regime_supported", "kvae_regime_sample_fwd", "kvae_regime_sample_bwd",
    "kvae_dp_last_error", "kvae_dp_handle_bytes", "kvae_dp_create", "kvae_dp_connect", "kvae_dp_destroy", "kvae_dp_finalize", "kvae_kf_bwd_dp",
    "kvae_vae_last_error", "kvae_vae_loss_workspace_bytes", "kvae_vae_loss_fwd", "kvae_vae_loss_bwd", "kvae_vae_reparam_fwd", "kvae_vae_reparam_bwd",
]


def _ptr(t, name, device=None):
    if t is None:
        return None
    p = t.data_ptr()
    # one combined test on the hot path; the specific message is worked out only when it fails
    if not (t.is_cuda and (t.dtype is torch.float32 or t.dtype is torch.int32) and t.is_contiguous() and p % 16 == 0
            and (device is None or t.device == device)):
        if not t.is_cuda:
            raise KvaeError(f"{name}: expected a CUDA tensor (the Kalman hot path has no CPU implementation)")
        if t.dtype != torch.float32 and t.dtype != torch.int32:
            raise KvaeError(f"{name}: expected float32, got {t.dtype}")
        if not t.is_contiguous():
            raise KvaeError(f"{name}: expected a contiguous tensor")
        if p % 16 != 0:
            raise KvaeError(f"{name}: base pointer must be 16-byte aligned")
        raise KvaeError(f"{name}: on {t.device}, expected {device}")
    return p


def _check(rc, what, L=None):
    if rc != 0:
        raise KvaeError(f"{what} failed (status {rc}): {(L or lib()).kvae_last_error().decode()}")


FLAG_SMOOTH_ONLY = 1
FLAG_ELBO_ONLY = 2
FLAG_WITH_ELBO = 4     # kvae_kf_bwd evaluates the ELBO too (writes terms) and normalises the gradients itself
FLAG_RAW_SUMS = 8      # ... but leaves the 1/max(sum mask,1) factor to the caller (data parallel)


def make_dims(B, T, n, p, m, K, q_per_mode, c_shared, lanes=0, flags=0):
    return KvaeDims(B, T, n, p, m, K, int(bool(q_per_mode)), int(bool(c_shared)), int(lanes), int(flags))


def supported(dims) -> bool:
    return bool(lib_for(dims).kvae_supported(byref(dims)))


def pick_lanes(dims) -> int:
    return int(lib_for(dims).kvae_pick_lanes(byref(dims)))


def make_inputs(Y, U, mask, alpha, A, Bm, C, Q, R, mu0, Sigma0, mu_init=None, Sigma_init=None,
                A_dense=None, B_dense=None, C_dense=None, Q_dense=None):
    dev = Y.device
    names = ("Y", "U", "mask", "alpha", "A", "Bm", "C", "Q", "R", "mu0", "Sigma0", "mu_init", "Sigma_init",
             "A_dense", "B_dense", "C_dense", "Q_dense")
    vals = (Y, U, mask, alpha, A, Bm, C, Q, R, mu0, Sigma0, mu_init, Sigma_init, A_dense, B_dense, C_dense, Q_dense)
    return KvaeInputs(*[_ptr(v, k, dev) for k, v in zip(names, vals)])


def mask_partials_count(dims) -> int:
    return int(lib_for(dims).kvae_kf_mask_partials_count(byref(dims)))


def make_states(mus_filt, Sigmas_filt, mus_pred, Sigmas_pred, mus_smooth=None, Sigmas_smooth=None, mask_partials=None,
                a_filt=None, a_smooth=None):
    names = ("mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred", "mus_smooth", "Sigmas_smooth", "mask_partials", "a_filt", "a_smooth")
    vals = (mus_filt, Sigmas_filt, mus_pred, Sigmas_pred, mus_smooth, Sigmas_smooth, mask_partials, a_filt, a_smooth)
    return KvaeStates(*[_ptr(v, k) for k, v in zip(names, vals)])


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(device):
    """cudaStream_t of torch's current stream on `device` (the raw getter skips building a torch.cuda.Stream object)."""
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(device).cuda_stream


def filter_smooth_fwd(dims, inputs, states, A_list, B_list, C_list, info, device):
    rc = lib_for(dims).kvae_kf_filter_smooth_fwd(byref(dims), byref(inputs), byref(states), _ptr(A_list, "A_list"),
                                         _ptr(B_list, "B_list"), _ptr(C_list, "C_list"), _ptr(info, "info"),
                                         device.index, _stream(device))
    _check(rc, "kvae_kf_filter_smooth_fwd", lib_for(dims))


def filter_lstm_fwd(dims, inputs, states, A_list, B_list, C_list, lstm_tensors, hidden, alpha_out, info, device):
    """Filter sweep with the LSTM dynamics network in the loop.  lstm_tensors: dict with w_ih, w_hh, b_ih, b_hh, w_head,
    b_head (+ optional h0, c0, h_out, c_out)."""
    ls = KvaeLstm(*[_ptr(lstm_tensors.get(k), "lstm." + k) for k, _ in KvaeLstm._fields_[:-1]], int(hidden))
    rc = lib_for(dims).kvae_kf_filter_lstm_fwd(byref(dims), byref(inputs), byref(states), _ptr(A_list, "A_list"), _ptr(B_list, "B_list"),
                                       _ptr(C_list, "C_list"), byref(ls), _ptr(alpha_out, "alpha_out"), _ptr(info, "info"),
                                       device.index, _stream(device))
    _check(rc, "kvae_kf_filter_lstm_fwd", lib_for(dims))


def elbo_workspace_bytes(dims) -> int:
    return int(lib_for(dims).kvae_kf_elbo_workspace_bytes(byref(dims)))


def _jit(jitter):
    """jitter: a float (both factorisations, no fallback) or the rung of the reference's _safe_cholesky ladder as a
    tuple (jitter_smooth, jitter_q, diag_smooth, diag_q) -> (float, KvaeCholOpts | None)"""
    if isinstance(jitter, tuple):
        js, jq, ds, dq = jitter
        return float(js), KvaeCholOpts(float(jq), int(bool(ds)), int(bool(dq)))
    return float(jitter), None


def elbo_fwd(dims, inputs, states, eps, jitter, terms, workspace, info, device):
    js, opts = _jit(jitter)
    rc = lib_for(dims).kvae_kf_elbo_fwd_ex(byref(dims), byref(inputs), byref(states), _ptr(eps, "eps"), c_float(js),
                                   byref(opts) if opts is not None else None,
                                   _ptr(terms, "terms"), c_void_p(workspace.data_ptr()), _ptr(info, "info"),
                                   device.index, _stream(device))
    _check(rc, "kvae_kf_elbo_fwd", lib_for(dims))


def bwd_workspace_bytes(dims) -> int:
    return int(lib_for(dims).kvae_kf_bwd_workspace_bytes(byref(dims)))


def bwd(dims, inputs, states, eps, jitter, g_elbo, terms, cot, grads, workspace, info, device):
    cot_s = KvaeCotangents(*[_ptr(cot.get(k) if cot else None, "cot." + k) for k, _ in KvaeCotangents._fields_])
    grads_s = KvaeGrads(*[_ptr(grads.get(k), k) for k, _ in KvaeGrads._fields_])
    js, opts = _jit(jitter)
    rc = lib_for(dims).kvae_kf_bwd_ex(byref(dims), byref(inputs), byref(states), _ptr(eps, "eps"), c_float(js),
                              byref(opts) if opts is not None else None,
                              _ptr(g_elbo, "g_elbo"), _ptr(terms, "terms"), byref(cot_s), byref(grads_s),
                              c_void_p(workspace.data_ptr()), _ptr(info, "info"), device.index, _stream(device))
    _check(rc, "kvae_kf_bwd", lib_for(dims))


# ---------------------------------------------------------------------------------------------------------
# data-parallel exchange over NVLink peer memory (kvae_dp_*)
# ---------------------------------------------------------------------------------------------------------
def _check_dp(rc, what):
    if rc != 0:
        raise KvaeError(f"{what} failed (status {rc}): {lib().kvae_dp_last_error().decode()}")


def dp_create(device, rank, world, nfloats):
    """-> (comm handle (c_void_p), IPC handle bytes to all-gather)"""
    L = lib()
    comm = c_void_p()
    buf = ctypes.create_string_buffer(int(L.kvae_dp_handle_bytes()))
    _check_dp(L.kvae_dp_create(device.index, rank, world, nfloats, byref(comm), buf), "kvae_dp_create")
    return comm, buf.raw


def dp_connect(comm, handles):
    blob = b"".join(handles)
    _check_dp(lib().kvae_dp_connect(comm, ctypes.c_char_p(blob)), "kvae_dp_connect")


def dp_destroy(comm):
    if comm:
        lib().kvae_dp_destroy(comm)


def bwd_dp(dims, inputs, states, eps, jitter, g_elbo, terms, grads, workspace, info, device, comm):
    """kvae_kf_bwd + the cross-rank exchange in its final kernel (dims.flags must hold WITH_ELBO | RAW_SUMS)."""
    grads_s = KvaeGrads(*[_ptr(grads.get(k), k) for k, _ in KvaeGrads._fields_])
    rc = lib_for(dims).kvae_kf_bwd_dp(byref(dims), byref(inputs), byref(states), _ptr(eps, "eps"), c_float(jitter),
                              _ptr(g_elbo, "g_elbo"), _ptr(terms, "terms"), byref(grads_s), c_void_p(workspace.data_ptr()),
                              _ptr(info, "info"), device.index, _stream(device), comm)
    _check(rc, "kvae_kf_bwd_dp", lib_for(dims))


def dp_finalize(dims, comm, grads, terms, info, device):
    grads_s = KvaeGrads(*[_ptr(grads.get(k), k) for k, _ in KvaeGrads._fields_])
    rc = lib().kvae_dp_finalize(byref(dims), comm, byref(grads_s), _ptr(terms, "terms"), _ptr(info, "info"), _stream(device))
    _check_dp(rc, "kvae_dp_finalize")


# ---------------------------------------------------------------------------------------------------------
# SKVAE regime sampler (kvae_regime_*)
# ---------------------------------------------------------------------------------------------------------
def _check_rg(rc, what):
    if rc != 0:
        raise KvaeError(f"{what} failed (status {rc}): {lib().kvae_regime_last_error().decode()}")


def regime_fwd(B, T, K, hard, tau, logits, init_logits, gumbel, trans, y_seq, log_q, log_p, device):
    d = KvaeRegimeDims(B, T, K, int(bool(hard)), float(tau))
    rc = lib().kvae_regime_sample_fwd(byref(d), _ptr(logits, "logits"), _ptr(init_logits, "init_logits"), _ptr(gumbel, "gumbel"),
                                      _ptr(trans, "trans"), _ptr(y_seq, "y_seq"), _ptr(log_q, "log_q"), _ptr(log_p, "log_p"),
                                      device.index, _stream(device))
    _check_rg(rc, "kvae_regime_sample_fwd")


def regime_bwd(B, T, K, hard, tau, logits, init_logits, gumbel, trans, y_seq, g_y, g_logq, g_logp, d_logits, d_init, device):
    d = KvaeRegimeDims(B, T, K, int(bool(hard)), float(tau))
    rc = lib().kvae_regime_sample_bwd(byref(d), _ptr(logits, "logits"), _ptr(init_logits, "init_logits"), _ptr(gumbel, "gumbel"),
                                      _ptr(trans, "trans"), _ptr(y_seq, "y_seq"), _ptr(g_y, "g_y"), _ptr(g_logq, "g_logq"),
                                      _ptr(g_logp, "g_logp"), _ptr(d_logits, "d_logits"), _ptr(d_init, "d_init"),
                                      device.index, _stream(device))
    _check_rg(rc, "kvae_regime_sample_bwd")
